"""Shared test fixtures: tiny models, synthetic clients."""

from __future__ import annotations

from pathlib import Path

import torch
from torch import nn

from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.common.typing import Config
from fl4health_b200.engine.data import BatchedTensorLoader
from fl4health_b200.engine.options import EngineOptions
from fl4health_b200.metrics import Accuracy
from fl4health_b200.models import Net
from fl4health_b200.utils.dataset import TensorDataset


def synthetic_cifar(n: int, seed: int, num_classes: int = 10) -> TensorDataset:
    gen = torch.Generator().manual_seed(seed)
    targets = torch.randint(0, num_classes, (n,), generator=gen)
    # class-dependent means so that a few SGD steps measurably reduce the loss
    data = torch.randn(n, 3, 32, 32, generator=gen) * 0.5 + (targets.float().view(-1, 1, 1, 1) - 4.5) * 0.2
    return TensorDataset(data, targets)


class TinyNet(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.conv = nn.Conv2d(3, 4, 3, padding=1)
        self.bn = nn.BatchNorm2d(4)
        self.fc = nn.Linear(4 * 8 * 8, 10)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = torch.nn.functional.adaptive_avg_pool2d(torch.relu(self.bn(self.conv(x))), 8)
        return self.fc(torch.flatten(x, 1))


class SyntheticCifarMixin:
    """User hooks on synthetic CIFAR-shaped data; combine with any client class: ``class C(SyntheticCifarMixin, X)``.
    Knobs are plain attributes so the mixin never touches ``__init__`` signatures."""

    seed: int = 0
    n_train: int = 256
    n_val: int = 64
    model_fn = staticmethod(Net)
    lr: float = 0.05
    momentum: float = 0.9
    optimizer_name: str = "sgd"

    def get_model(self, config: Config) -> nn.Module:
        torch.manual_seed(1234)  # same init everywhere (the server overrides it anyway)
        return self.model_fn()

    def get_data_loaders(self, config: Config):  # noqa: ANN201
        bs = int(config.get("batch_size", 32))
        train = BatchedTensorLoader(synthetic_cifar(self.n_train, self.seed), bs, shuffle=True,
                                    generator=torch.Generator().manual_seed(self.seed))
        val = BatchedTensorLoader(synthetic_cifar(self.n_val, 10_000 + self.seed), bs)
        return train, val

    def get_criterion(self, config: Config):  # noqa: ANN201
        return nn.CrossEntropyLoss()

    def get_optimizer(self, config: Config):  # noqa: ANN201
        if self.optimizer_name == "adamw":
            return torch.optim.AdamW(self.model.parameters(), lr=self.lr)
        return torch.optim.SGD(self.model.parameters(), lr=self.lr, momentum=self.momentum)


class SyntheticCifarClient(SyntheticCifarMixin, BasicClient):
    def __init__(self, *args, seed: int = 0, n_train: int = 256, n_val: int = 64, model_fn=Net, lr: float = 0.05,  # noqa: ANN001, ANN002
                 momentum: float = 0.9, optimizer: str = "sgd", **kwargs) -> None:  # noqa: ANN003
        super().__init__(*args, **kwargs)
        self.seed, self.n_train, self.n_val, self.model_fn = seed, n_train, n_val, model_fn
        self.lr, self.momentum, self.optimizer_name = lr, momentum, optimizer


def make_mixed_clients(client_cls, k: int, device: str = "cpu", engine: EngineOptions | None = None, **attrs):  # noqa: ANN001, ANN003, ANN201
    """K clients of ``class _(SyntheticCifarMixin, client_cls)`` with per-client data seeds."""
    mixed = type(f"Synthetic{client_cls.__name__}", (SyntheticCifarMixin, client_cls), {})
    clients = []
    for idx in range(k):
        client = mixed(Path("."), [Accuracy()], torch.device(device), client_name=f"c{idx}", engine_options=engine)
        client.seed = idx
        for key, value in attrs.items():
            setattr(client, key, value)
        clients.append(client)
    return clients


def make_clients(k: int, device: str = "cpu", engine: EngineOptions | None = None, **kwargs) -> list[SyntheticCifarClient]:  # noqa: ANN003
    return [
        SyntheticCifarClient(Path("."), [Accuracy()], torch.device(device), client_name=f"c{idx}", seed=idx,
                             engine_options=engine, **kwargs)
        for idx in range(k)
    ]


def fit_config_fn(local_steps: int = 5, batch_size: int = 32):  # noqa: ANN201
    def fn(server_round: int) -> Config:
        return {"current_server_round": server_round, "local_steps": local_steps, "batch_size": batch_size}

    return fn
