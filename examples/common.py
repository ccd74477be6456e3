"""Shared pieces of the example scenarios: data (real MNIST / CIFAR-10 when present under ``data_dir``, synthetic
tensors of the same shape otherwise — the box running the smoke tests has no datasets), a client mixin that
implements the data / loss / optimizer hooks from the YAML config, and strategy keyword helpers."""

from __future__ import annotations

from collections.abc import Callable
from pathlib import Path
from typing import Any

import torch
import yaml
from torch import nn

from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config
from fl4health_b200.engine.data import BatchedTensorLoader
from fl4health_b200.metrics.metric_aggregation import evaluate_metrics_aggregation_fn, fit_metrics_aggregation_fn
from fl4health_b200.utils.dataset import TensorDataset
from fl4health_b200.utils.sampler import DirichletLabelBasedSampler

CONFIG_DIR = Path(__file__).resolve().parent / "configs"
DEFAULTS: dict[str, Any] = {
    "n_server_rounds": 2, "n_clients": 2, "batch_size": 32, "local_steps": 4, "dataset": "mnist", "data_dir": "examples/datasets",
    "samples_per_client": 256, "val_samples_per_client": 64, "learning_rate": 0.05, "optimizer": "sgd", "momentum": 0.0,
    "heterogeneity_beta": None, "seed": 2024,
}
SHAPES = {"mnist": ((1, 28, 28), 10), "cifar": ((3, 32, 32), 10)}


# The reference's example configs name the same knobs differently per example; a YAML written for it can be passed to
# ``--config`` as is.  Left: the reference's key; right: the key the scenarios read.
REFERENCE_KEY_ALIASES = {
    "initial_proximal_weight": "initial_loss_weight", "adapt_proximal_weight": "adapt_loss_weight",
    "adaptive_proximal_weight": "adapt_loss_weight", "proximal_weight_delta": "loss_weight_delta",
    "proximal_weight_patience": "loss_weight_patience", "beta": "heterogeneity_beta",
    "server_noise_multiplier": "weight_noise_multiplier", "clipping_bit_noise_multiplier": "clipping_noise_multiplier",
    "clipping_bound": "initial_clipping_bound", "server_momentum": "server_beta", "weighted_averaging": "weighted_aggregation",
    "ckpt_path": "checkpoint_path",
}


def translate_reference_keys(raw: dict[str, Any]) -> dict[str, Any]:
    """Reference-style keys renamed to the scenarios' names (a key given under both names keeps the native one); a
    null ``local_epochs`` / ``local_steps`` -- the reference writes the unused one as null -- is dropped."""
    out = {key: value for key, value in raw.items() if key not in REFERENCE_KEY_ALIASES}
    for key, value in raw.items():
        if key in REFERENCE_KEY_ALIASES:
            out.setdefault(REFERENCE_KEY_ALIASES[key], value)
    for duration in ("local_epochs", "local_steps"):
        if duration in out and out[duration] is None:
            del out[duration]
    return out  # (when ``local_epochs`` is given it takes precedence over the default step count: make_config_fn)


def load_example_config(scenario: str, path: str | None = None, overrides: dict[str, Any] | None = None) -> dict[str, Any]:
    config = dict(DEFAULTS)
    file = Path(path) if path else CONFIG_DIR / f"{scenario}.yaml"
    if file.exists():
        config.update(translate_reference_keys(yaml.safe_load(file.read_text()) or {}))
    config.update({k: v for k, v in (overrides or {}).items() if v is not None})
    return config


def _synthetic(shape: tuple[int, ...], classes: int, n: int, seed: int) -> TensorDataset:
    gen = torch.Generator().manual_seed(seed)
    targets = torch.randint(0, classes, (n,), generator=gen)
    data = torch.randn(n, *shape, generator=gen) * 0.5 + (targets.float().view(-1, *([1] * len(shape))) - classes / 2) * 0.2
    return TensorDataset(data, targets)


def client_datasets(config: dict[str, Any], client_index: int) -> tuple[TensorDataset, TensorDataset]:
    """(train, validation) datasets of one client: a seeded shard of the real dataset when it is on disk, synthetic
    class-separable tensors otherwise."""
    name = config["dataset"]
    shape, classes = SHAPES[name]
    n_train, n_val = int(config["samples_per_client"]), int(config["val_samples_per_client"])
    try:
        from fl4health_b200.utils import load_data

        loader = load_data.get_train_and_val_mnist_datasets if name == "mnist" else load_data.get_train_and_val_cifar10_datasets
        train, val = loader(Path(config["data_dir"]), hash_key=config["seed"])
        gen = torch.Generator().manual_seed(config["seed"] + client_index)
        pick = torch.randperm(len(train), generator=gen)[:n_train]
        train.data, train.targets = train.data[pick], train.targets[pick]
        pick = torch.randperm(len(val), generator=gen)[:n_val]
        val.data, val.targets = val.data[pick], val.targets[pick]
    except (FileNotFoundError, OSError, RuntimeError):
        train = _synthetic(shape, classes, n_train, config["seed"] + 17 * client_index)
        val = _synthetic(shape, classes, n_val, config["seed"] + 10_000 + client_index)
    beta = config.get("heterogeneity_beta")
    if beta:
        sampler = DirichletLabelBasedSampler(list(range(classes)), hash_key=config["seed"] + client_index, sample_percentage=0.75, beta=float(beta))
        train = sampler.subsample(train)
    return train, val


def _rng_devices() -> list[int]:
    """Devices whose generators ``torch.manual_seed`` would rewind: forked together with the CPU generator so that a
    seeded model construction leaves the client's random streams (mask sampling, dropout, DP noise) where they were."""
    return [torch.cuda.current_device()] if torch.cuda.is_available() and torch.cuda.is_initialized() else []


class ExampleClientMixin:
    """Hooks shared by every example client (combine as ``class C(ExampleClientMixin, SomeClient)``).  The scenario
    sets ``example_config``, ``client_index`` and ``model_factory`` on the instance."""

    example_config: dict[str, Any]
    client_index: int = 0
    model_factory: Callable[[], nn.Module]

    def get_model(self, config: Config) -> nn.Module:
        # every client starts from the same initialisation, without rewinding the process' random stream (which the
        # client keeps using for mask sampling, dropout, DP noise ...)
        with torch.random.fork_rng(devices=_rng_devices()):
            torch.manual_seed(self.example_config["seed"])
            return self.model_factory()

    def get_data_loaders(self, config: Config) -> tuple[BatchedTensorLoader, BatchedTensorLoader]:
        train, val = client_datasets(self.example_config, self.client_index)
        batch_size = int(config.get("batch_size", self.example_config["batch_size"]))
        placement = "device" if self.device.type == "cuda" else "host"  # type: ignore[attr-defined]
        gen = torch.Generator().manual_seed(self.example_config["seed"] + self.client_index)
        return (BatchedTensorLoader(train, batch_size, shuffle=True, placement=placement, device=self.device, generator=gen),  # type: ignore[attr-defined]
                BatchedTensorLoader(val, batch_size, placement=placement, device=self.device))  # type: ignore[attr-defined]

    def get_criterion(self, config: Config) -> nn.Module:
        return nn.CrossEntropyLoss()

    def make_optimizer(self, params: Any) -> torch.optim.Optimizer:
        cfg = self.example_config
        if cfg["optimizer"] == "adamw":
            return torch.optim.AdamW(params, lr=cfg["learning_rate"])
        return torch.optim.SGD(params, lr=cfg["learning_rate"], momentum=cfg["momentum"])

    def get_optimizer(self, config: Config) -> Any:
        return self.make_optimizer(self.model.parameters())  # type: ignore[attr-defined]


def make_config_fn(config: dict[str, Any], **extra: Any) -> Callable[[int], Config]:
    keys = ("batch_size", "local_steps", "local_epochs", "n_server_rounds")

    def fn(server_round: int) -> Config:
        out: Config = {"current_server_round": server_round, **{k: config[k] for k in keys if config.get(k) is not None}, **extra}
        if "local_epochs" in out:
            out.pop("local_steps", None)
        return out

    return fn


def strategy_kwargs(config: dict[str, Any], config_fn: Callable[[int], Config] | None = None) -> dict[str, Any]:
    n = int(config["n_clients"])
    fn = config_fn or make_config_fn(config)
    return dict(min_fit_clients=n, min_evaluate_clients=n, min_available_clients=n, on_fit_config_fn=fn, on_evaluate_config_fn=fn,
                fit_metrics_aggregation_fn=fit_metrics_aggregation_fn, evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn)


def describe(scenario: str, config: dict[str, Any]) -> None:
    from logging import INFO

    log(INFO, f"[example:{scenario}] " + ", ".join(f"{k}={v}" for k, v in sorted(config.items())))
