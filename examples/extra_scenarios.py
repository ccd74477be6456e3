"""Scenarios that need more than a client class + a model: evaluation-only federations, one-shot model merging,
nnU-Net (with an injectable toy segmentation backend when ``nnunetv2`` is not installed), LoRA-adapter federation,
the MMD-regularised personalised clients, conditional VAEs and weighted client-level DP on tabular data
(SURVEY Appendix C rows ``federated_eval_example``, ``model_merge_example``, ``nnunet_example``, ``nnunet_pfl_example``,
``fedllm_example``, ``mr_mtl_*``/``ditto_*`` MMD variants, ``ae_examples/cvae_*``, ``client_level_dp_weighted``)."""

from __future__ import annotations

import math
from pathlib import Path
from typing import Any

import torch
from torch import nn

from examples.common import ExampleClientMixin, client_datasets, strategy_kwargs
from examples.models import SmallCnn
from examples.scenarios import _adaptive_strategy, _dict_optimizers, _dp_fn, _fl_server, make_clients, scenario
from fl4health_b200.engine.data import BatchedTensorLoader
from fl4health_b200.metrics import Accuracy
from fl4health_b200.metrics.metric_aggregation import evaluate_metrics_aggregation_fn, fit_metrics_aggregation_fn
from fl4health_b200.servers.client_manager import SimpleClientManager
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg


# ---------------------------------------------------------------------------------------------------------------
# evaluation-only and one-shot merge federations
# ---------------------------------------------------------------------------------------------------------------
@scenario("federated_eval_example")
def federated_eval_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    """Every client evaluates a server-provided global checkpoint and its own local checkpoint on local data."""
    from fl4health_b200.clients.evaluate_client import EvaluateClient
    from fl4health_b200.servers.evaluate_server import EvaluateServer

    out = Path(config.get("checkpoint_dir", "examples_out/federated_eval"))
    out.mkdir(parents=True, exist_ok=True)
    torch.manual_seed(config["seed"])
    torch.save(SmallCnn(config["dataset"]), out / "global.pkl")

    class Client(ExampleClientMixin, EvaluateClient):
        def get_data_loader(self, cfg: dict[str, Any]) -> tuple[BatchedTensorLoader]:
            _, val = client_datasets(self.example_config, self.client_index)
            return (BatchedTensorLoader(val, self.example_config["batch_size"]),)

        def initialize_global_model(self, cfg: dict[str, Any]) -> nn.Module:
            return SmallCnn(self.example_config["dataset"])

    clients = []
    for index in range(int(config["n_clients"])):
        torch.manual_seed(config["seed"] + 1 + index)
        torch.save(SmallCnn(config["dataset"]), out / f"local_{index}.pkl")
        client = Client(Path(config["data_dir"]), [Accuracy()], device, model_checkpoint_path=out / f"local_{index}.pkl",
                        client_name=f"client_{index}")
        client.example_config, client.client_index = config, index
        clients.append(client)
    server = EvaluateServer(SimpleClientManager(), fraction_evaluate=1.0, model_checkpoint_path=out / "global.pkl",
                            evaluate_config={"current_server_round": 0}, min_available_clients=int(config["n_clients"]),
                            evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn)
    return server, clients


@scenario("model_merge_example")
def model_merge_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    """Clients ship locally pre-trained weights once; the server averages them, evaluates and checkpoints the merge."""
    from fl4health_b200.checkpointing.checkpointer import LatestTorchModuleCheckpointer
    from fl4health_b200.clients.model_merge_client import ModelMergeClient
    from fl4health_b200.parameter_exchange.full_exchanger import FullParameterExchanger
    from fl4health_b200.servers.model_merge_server import ModelMergeServer
    from fl4health_b200.strategies.model_merge_strategy import ModelMergeStrategy

    out = Path(config.get("checkpoint_dir", "examples_out/model_merge"))
    out.mkdir(parents=True, exist_ok=True)

    class Client(ModelMergeClient):
        example_config: dict[str, Any]
        client_index = 0

        def get_model(self, cfg: dict[str, Any]) -> nn.Module:
            torch.manual_seed(self.example_config["seed"] + self.client_index)  # stands in for "pre-trained locally"
            return SmallCnn(self.example_config["dataset"])

        def get_test_data_loader(self, cfg: dict[str, Any]) -> BatchedTensorLoader:
            _, val = client_datasets(self.example_config, self.client_index)
            return BatchedTensorLoader(val, self.example_config["batch_size"])

    clients = []
    for index in range(int(config["n_clients"])):
        client = Client(Path(config["data_dir"]), out / f"client_{index}.pt", [Accuracy()], device, client_name=f"client_{index}")
        client.example_config, client.client_index = config, index
        clients.append(client)
    n = int(config["n_clients"])
    strategy = ModelMergeStrategy(min_fit_clients=n, min_evaluate_clients=n, min_available_clients=n,
                                  fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
                                  evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn, weighted_aggregation=False)
    server = ModelMergeServer(SimpleClientManager(), strategy, LatestTorchModuleCheckpointer(str(out), "merged.pkl"),
                              SmallCnn(config["dataset"]), FullParameterExchanger())
    return server, clients


# ---------------------------------------------------------------------------------------------------------------
# nnU-Net
# ---------------------------------------------------------------------------------------------------------------
class ToySegNet(nn.Module):
    """Two-resolution segmentation net with optional deep supervision (list output in training mode)."""

    def __init__(self, in_channels: int, heads: int, deep_supervision: bool) -> None:
        super().__init__()
        self.body = nn.Sequential(nn.Conv2d(in_channels, 16, 3, padding=1), nn.ReLU(), nn.Conv2d(16, 16, 3, padding=1), nn.ReLU())
        self.head, self.low_head = nn.Conv2d(16, heads, 1), nn.Conv2d(16, heads, 1)
        self.deep_supervision = deep_supervision

    def forward(self, x: torch.Tensor) -> Any:
        h = self.body(x)
        if self.deep_supervision and self.training:
            return [self.head(h), self.low_head(nn.functional.avg_pool2d(h, 2))]
        return self.head(h)


class _DeepSupervisionCe(nn.Module):
    def forward(self, preds: Any, targets: Any) -> torch.Tensor:
        if isinstance(preds, (list, tuple)):
            return sum(w * nn.functional.cross_entropy(p, t[:, 0].long()) for w, p, t in zip((1.0, 0.5), preds, targets))  # type: ignore[return-value]
        return nn.functional.cross_entropy(preds, targets[:, 0].long())


class _BlobBatches:
    """Infinite stream of ``{"data", "target"}`` batches shaped like nnU-Net's augmenter output: 2-channel 32×32
    images whose labels are thresholded sums of the channels (3 classes)."""

    def __init__(self, seed: int, deep_supervision: bool, batch_size: int) -> None:
        self.gen, self.deep_supervision, self.batch_size = torch.Generator().manual_seed(seed), deep_supervision, batch_size

    def __iter__(self) -> Any:
        while True:
            x = torch.randn(self.batch_size, 2, 32, 32, generator=self.gen)
            y = (x[:, :1] + x[:, 1:] > 0).long() + (x[:, :1] > 1).long()
            yield {"data": x, "target": [y, y[:, :, ::2, ::2]] if self.deep_supervision else y}


class SyntheticSegmentationBackend:
    """Stands in for ``Nnunetv2Backend`` when nnunetv2 / the MSD data are not available: same protocol (``plan`` →
    plans dict, ``prepare`` → ``PreparedExperiment``), synthetic blobs, a two-resolution CNN with deep supervision."""

    dataset_name = "Dataset900_SyntheticBlobs"

    def __init__(self, seed: int, steps_per_epoch: int = 4) -> None:
        self.seed, self.steps_per_epoch = seed, steps_per_epoch

    def plan(self) -> dict[str, Any]:
        return {"plans_name": "synthetic_plans", "dataset_name": self.dataset_name,
                "configurations": {"2d": {"median_image_size_in_voxels": [32, 32], "batch_size": 4, "patch_size": [32, 32]}}}

    def prepare(self, plans: dict[str, Any], config: Any, fold: Any, batch_size: int, device: torch.device) -> Any:
        from fl4health_b200.clients.nnunet_client import LabelInfo, PreparedExperiment
        from fl4health_b200.utils.nnunet_utils import NnUNetDataLoaderWrapper

        torch.manual_seed(0)
        return PreparedExperiment(
            network=ToySegNet(2, 3, True), loss=_DeepSupervisionCe(),
            train_loader=NnUNetDataLoaderWrapper(_BlobBatches(self.seed, True, batch_size), config, set_len=self.steps_per_epoch),
            val_loader=NnUNetDataLoaderWrapper(_BlobBatches(1000 + self.seed, False, batch_size), config, set_len=2),
            labels=LabelInfo(ignore_label=None, has_regions=False, num_segmentation_heads=3), num_input_channels=2,
            enable_deep_supervision=True, initial_lr=0.05,
        )


def _nnunet_backend(config: dict[str, Any], index: int) -> Any:
    if config.get("dataset_id") is not None:  # a real MSD dataset id: use nnunetv2 (must be installed, data preprocessed)
        return None
    return SyntheticSegmentationBackend(config["seed"] + index)


def _nnunet_federation(config: dict[str, Any], device: torch.device, client_cls: type, strategy_factory: Any = None,
                       **client_kwargs: Any) -> tuple[Any, list[Any]]:
    from fl4health_b200.metrics.efficient_metrics import MultiClassDice
    from fl4health_b200.servers.nnunet_server import NnunetServer

    def fn(server_round: int) -> dict[str, Any]:
        return {"current_server_round": server_round, "local_epochs": int(config.get("local_epochs", 1)), "batch_size": 4,
                "nnunet_config": config.get("nnunet_config", "2d"), "n_server_rounds": config["n_server_rounds"]}

    clients = [client_cls(device, int(config.get("dataset_id") or 900), fold=0,
                          metrics=[MultiClassDice(batch_dim=None, label_dim=1, threshold=1)], backend=_nnunet_backend(config, i),
                          client_name=f"client_{i}", verbose=False, **client_kwargs) for i in range(int(config["n_clients"]))]
    strategy = strategy_factory(fn) if strategy_factory is not None else BasicFedAvg(**strategy_kwargs(config, fn))
    server = NnunetServer(SimpleClientManager(), {"n_server_rounds": config["n_server_rounds"], "nnunet_config": "2d"}, fn, strategy,
                          model_builder=lambda plans, cfg, in_ch, heads, ds: ToySegNet(in_ch, heads, ds))
    return server, clients


@scenario("nnunet_example")
def nnunet_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.nnunet_client import NnunetClient

    return _nnunet_federation(config, device, NnunetClient)


@scenario("nnunet_pfl_example")
def nnunet_pfl_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    """Personalised nnU-Net: ``make_it_personal(FlexibleNnunetClient, DITTO)`` — a personal segmentation model trained
    next to the federated one with a drift penalty."""
    from fl4health_b200.clients.flexible.nnunet import FlexibleNnunetClient
    from fl4health_b200.mixins.personalized import PersonalizedMode, make_it_personal

    from fl4health_b200.strategies.fedavg_with_adaptive_constraint import FedAvgWithAdaptiveConstraint

    mode = PersonalizedMode(config.get("personalized_strategy", "ditto"))

    def strategy(fn: Any) -> Any:  # the drift-penalty weight rides with the parameters (adaptive-constraint packing)
        return FedAvgWithAdaptiveConstraint(initial_parameters=None, initial_loss_weight=0.1, adapt_loss_weight=False,
                                            **strategy_kwargs(config, fn))

    return _nnunet_federation(config, device, make_it_personal(FlexibleNnunetClient, mode), strategy)


# ---------------------------------------------------------------------------------------------------------------
# LoRA adapters (fedllm_example): only the adapter tensors travel
# ---------------------------------------------------------------------------------------------------------------
class LoraLinear(nn.Module):
    """``y = W x + (alpha / r) · B A x`` with ``W`` frozen; parameter names carry the ``lora_`` marker PEFT uses so
    ``utils/peft_parameter_extraction.get_peft_state_dict`` finds them."""

    def __init__(self, in_features: int, out_features: int, rank: int = 4, alpha: float = 8.0) -> None:
        super().__init__()
        self.base = nn.Linear(in_features, out_features)
        self.base.requires_grad_(False)
        self.lora_A = nn.Parameter(torch.randn(rank, in_features) / math.sqrt(in_features))
        self.lora_B = nn.Parameter(torch.zeros(out_features, rank))
        self.scale = alpha / rank

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.base(x) + (x @ self.lora_A.t() @ self.lora_B.t()) * self.scale


class TinyCausalLm(nn.Module):
    """A small decoder-only LM (embedding → N pre-LN blocks with LoRA on q/v and the MLP → tied-free LM head)."""

    def __init__(self, vocab: int = 256, width: int = 64, layers: int = 2, heads: int = 4, max_len: int = 64) -> None:
        super().__init__()
        self.embed, self.pos = nn.Embedding(vocab, width), nn.Embedding(max_len, width)
        self.embed.requires_grad_(False)
        self.pos.requires_grad_(False)
        self.blocks = nn.ModuleList(nn.ModuleDict({
            "ln1": nn.LayerNorm(width), "q": LoraLinear(width, width), "k": nn.Linear(width, width), "v": LoraLinear(width, width),
            "o": nn.Linear(width, width), "ln2": nn.LayerNorm(width), "up": LoraLinear(width, 4 * width), "down": nn.Linear(4 * width, width),
        }) for _ in range(layers))
        for block in self.blocks:
            for name in ("k", "o", "down", "ln1", "ln2"):
                block[name].requires_grad_(False)
        self.norm, self.lm_head, self.heads = nn.LayerNorm(width), nn.Linear(width, vocab, bias=False), heads
        self.norm.requires_grad_(False)
        self.lm_head.requires_grad_(False)

    def forward(self, tokens: torch.Tensor) -> torch.Tensor:
        b, t = tokens.shape
        h = self.embed(tokens) + self.pos(torch.arange(t, device=tokens.device))
        for block in self.blocks:
            x = block["ln1"](h)
            q, k, v = (block[n](x).view(b, t, self.heads, -1).transpose(1, 2) for n in ("q", "k", "v"))
            attn = nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(b, t, -1)
            h = h + block["o"](attn)
            h = h + block["down"](nn.functional.gelu(block["up"](block["ln2"](h))))
        return self.lm_head(self.norm(h)).transpose(1, 2)  # [B, vocab, T] for CrossEntropyLoss against [B, T]


@scenario("fedllm_example")
def fedllm_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    """Federated LoRA fine-tuning of a causal LM: the frozen base never leaves the client, FedAvg runs over the adapter
    tensors only (``FixedLayerExchanger`` over the PEFT state-dict keys)."""
    from fl4health_b200.clients.basic_client import BasicClient
    from fl4health_b200.parameter_exchange.layer_exchanger import FixedLayerExchanger
    from fl4health_b200.utils.dataset import TensorDataset
    from fl4health_b200.utils.peft_parameter_extraction import get_all_peft_parameters_from_model, get_peft_state_dict

    vocab, seq_len = int(config.get("vocab_size", 256)), int(config.get("seq_len", 32))

    def model_factory() -> nn.Module:
        return TinyCausalLm(vocab=vocab, max_len=seq_len)

    def shard(index: int, n: int, seed: int) -> TensorDataset:
        """Next-token data with learnable structure: token_{t+1} = (token_t * 3 + client-specific offset) mod vocab."""
        gen = torch.Generator().manual_seed(seed)
        start = torch.randint(0, vocab, (n, 1), generator=gen)
        steps = [start]
        for _ in range(seq_len):
            steps.append((steps[-1] * 3 + 1 + index) % vocab)
        tokens = torch.cat(steps, dim=1)
        return TensorDataset(tokens[:, :-1], tokens[:, 1:])

    def customise(client: Any) -> None:
        def loaders(cfg: dict[str, Any]) -> tuple[Any, Any]:
            return (BatchedTensorLoader(shard(client.client_index, config["samples_per_client"], config["seed"] + client.client_index),
                                        config["batch_size"], shuffle=True),
                    BatchedTensorLoader(shard(client.client_index, config["val_samples_per_client"], 10_000 + client.client_index),
                                        config["batch_size"]))

        client.get_data_loaders = loaders
        client.get_parameter_exchanger = lambda cfg: FixedLayerExchanger(list(get_peft_state_dict(client.model).keys()))
        client.get_optimizer = lambda cfg: client.make_optimizer([p for p in client.model.parameters() if p.requires_grad])

    clients = make_clients(BasicClient, config, device, model_factory, customise, metrics=[])
    torch.manual_seed(config["seed"])
    strategy = BasicFedAvg(initial_parameters=get_all_peft_parameters_from_model(model_factory()), **strategy_kwargs(config))
    return _fl_server(config, strategy), clients


# ---------------------------------------------------------------------------------------------------------------
# MMD-regularised personalised clients
# ---------------------------------------------------------------------------------------------------------------
def _mmd_scenario(config: dict[str, Any], device: torch.device, client_cls: type, server_cls: type, ditto: bool, **kwargs: Any) -> tuple[Any, list[Any]]:
    def customise(client: Any) -> None:
        if ditto:
            _dict_optimizers(client, {"global": lambda c: c.global_model, "local": lambda c: c.model})

    clients = make_clients(client_cls, config, device, lambda: SmallCnn(config["dataset"]), customise, **kwargs)
    return _fl_server(config, _adaptive_strategy({**config, "adapt_loss_weight": False}), server_cls), clients


def _feature_width(config: dict[str, Any], layer: str = "features") -> int:
    """Flattened width of the activations the feature-extractor hook captures for ``layer`` (prefix-matched like the
    clients do: the LAST module whose name starts with the prefix)."""
    from fl4health_b200.model_bases.feature_extractor_buffer import FeatureExtractorBuffer

    model = SmallCnn(config["dataset"])
    buffer = FeatureExtractorBuffer(model, {layer: True})
    buffer._maybe_register_hooks()
    shape = (1, 28, 28) if config["dataset"] == "mnist" else (3, 32, 32)
    with torch.no_grad():
        model(torch.zeros(2, *shape))
    return int(buffer.get_extracted_features()[layer].shape[1])


@scenario("mr_mtl_mkmmd_example")
def mr_mtl_mkmmd_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.mkmmd_clients import MrMtlMkMmdClient
    from fl4health_b200.servers.adaptive_constraint_servers.mrmtl_server import MrMtlServer

    return _mmd_scenario(config, device, MrMtlMkMmdClient, MrMtlServer, False, mkmmd_loss_weight=1.0,
                         feature_extraction_layers=["features"], beta_global_update_interval=2, num_accumulating_batches=2)


@scenario("ditto_deep_mmd_example")
def ditto_deep_mmd_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.deep_mmd_clients import DittoDeepMmdClient
    from fl4health_b200.servers.adaptive_constraint_servers.ditto_server import DittoServer

    return _mmd_scenario(config, device, DittoDeepMmdClient, DittoServer, True, deep_mmd_loss_weight=1.0,
                         feature_extraction_layers_with_size={"features": _feature_width(config)}, mmd_kernel_train_interval=2,
                         num_accumulating_batches=2)


@scenario("mr_mtl_deep_mmd_example")
def mr_mtl_deep_mmd_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.deep_mmd_clients import MrMtlDeepMmdClient
    from fl4health_b200.servers.adaptive_constraint_servers.mrmtl_server import MrMtlServer

    return _mmd_scenario(config, device, MrMtlDeepMmdClient, MrMtlServer, False, deep_mmd_loss_weight=1.0,
                         feature_extraction_layers_with_size={"features": _feature_width(config)}, mmd_kernel_train_interval=2,
                         num_accumulating_batches=2)


# ---------------------------------------------------------------------------------------------------------------
# conditional VAE (ae_examples/cvae_examples): the label is the condition, packed into the input by the converter
# ---------------------------------------------------------------------------------------------------------------
@scenario("cvae_example")
def cvae_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.basic_client import BasicClient
    from fl4health_b200.model_bases.autoencoders_base import ConditionalVae
    from fl4health_b200.preprocessing.autoencoders.loss import VaeLoss
    from fl4health_b200.utils.dataset_converter import AutoEncoderDatasetConverter

    in_dim, latent, classes = 28 * 28 if config["dataset"] == "mnist" else 3 * 32 * 32, 16, 10

    class Encoder(nn.Module):
        def __init__(self) -> None:
            super().__init__()
            self.body = nn.Sequential(nn.Linear(in_dim + classes, 64), nn.ReLU())
            self.mu, self.logvar = nn.Linear(64, latent), nn.Linear(64, latent)

        def forward(self, x: torch.Tensor, condition: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
            h = self.body(torch.cat((x.flatten(1), condition), dim=1))
            return self.mu(h), self.logvar(h)

    class Decoder(nn.Module):
        def __init__(self) -> None:
            super().__init__()
            self.body = nn.Sequential(nn.Linear(latent + classes, 64), nn.ReLU(), nn.Linear(64, in_dim))

        def forward(self, z: torch.Tensor, condition: torch.Tensor) -> torch.Tensor:
            return self.body(torch.cat((z, condition), dim=1))

    def customise(client: Any) -> None:
        converter = AutoEncoderDatasetConverter(condition="label", do_one_hot_encoding=True)

        def loaders(cfg: dict[str, Any]) -> tuple[Any, Any]:
            train, val = client_datasets(config, client.client_index)
            train_set = converter.convert_dataset(train)
            val_converter = AutoEncoderDatasetConverter(condition="label", do_one_hot_encoding=True, condition_vector_size=classes)
            return (BatchedTensorLoader(train_set, config["batch_size"], shuffle=True),
                    BatchedTensorLoader(val_converter.convert_dataset(val), config["batch_size"]))

        def model(cfg: dict[str, Any]) -> nn.Module:
            torch.manual_seed(config["seed"])
            shape = torch.Size((1, 28, 28) if config["dataset"] == "mnist" else (3, 32, 32))
            unpack = lambda packed: AutoEncoderDatasetConverter.unpack_input_condition(packed, classes, shape)  # noqa: E731
            return ConditionalVae(Encoder(), Decoder(), unpack)

        client.get_data_loaders, client.get_model = loaders, model
        client.get_criterion = lambda cfg: _FlatTargetVaeLoss(latent)

    class _FlatTargetVaeLoss(VaeLoss):
        def __init__(self, latent_dim: int) -> None:
            super().__init__(latent_dim, nn.MSELoss(reduction="sum"))

        def forward(self, preds: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
            return super().forward(preds, target.flatten(1))

    clients = make_clients(BasicClient, config, device, lambda: nn.Identity(), customise, metrics=[])
    return _fl_server(config, BasicFedAvg(**strategy_kwargs(config))), clients


# ---------------------------------------------------------------------------------------------------------------
# weighted client-level DP on tabular data (dp_fed_examples/client_level_dp_weighted)
# ---------------------------------------------------------------------------------------------------------------
@scenario("client_level_dp_weighted_example")
def client_level_dp_weighted_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    """Logistic regression on a 30-feature tabular task (breast-cancer shaped) with unequal client sizes: the server
    weighs clipped updates by capped client sample counts (``weighted_aggregation=True``)."""
    from fl4health_b200.client_managers.poisson_sampling_manager import PoissonSamplingClientManager
    from fl4health_b200.clients.clipping_client import NumpyClippingClient
    from fl4health_b200.servers.client_level_dp_fed_avg_server import ClientLevelDPFedAvgServer
    from fl4health_b200.strategies.client_dp_fedavgm import ClientLevelDPFedAvgM
    from fl4health_b200.utils.dataset import TensorDataset

    features = 30

    def shard(index: int, n: int, seed: int) -> TensorDataset:
        gen = torch.Generator().manual_seed(seed)
        x = torch.randn(n, features, generator=gen)
        w = torch.linspace(-1.0, 1.0, features)
        return TensorDataset(x, (x @ w + 0.1 * torch.randn(n, generator=gen) > 0).long())

    def customise(client: Any) -> None:
        n_train = int(config["samples_per_client"]) * (1 + client.client_index)  # unequal shards → non-trivial weights

        def loaders(cfg: dict[str, Any]) -> tuple[Any, Any]:
            return (BatchedTensorLoader(shard(client.client_index, n_train, config["seed"] + client.client_index), config["batch_size"], shuffle=True),
                    BatchedTensorLoader(shard(client.client_index, config["val_samples_per_client"], 5_000 + client.client_index), config["batch_size"]))

        client.get_data_loaders = loaders

    fn = _dp_fn(config)
    kwargs = {k: v for k, v in strategy_kwargs(config, fn).items() if k not in ("min_fit_clients", "min_evaluate_clients")}
    strategy = ClientLevelDPFedAvgM(fraction_fit=1.0, fraction_evaluate=1.0, adaptive_clipping=False, initial_clipping_bound=1.0,
                                    weight_noise_multiplier=0.05, weighted_aggregation=True, per_client_example_cap=float(
                                        2 * int(config["samples_per_client"])), **kwargs)
    server = ClientLevelDPFedAvgServer(PoissonSamplingClientManager(), {"n_server_rounds": config["n_server_rounds"]}, strategy,
                                       server_noise_multiplier=0.05, num_server_rounds=config["n_server_rounds"],
                                       on_init_parameters_config_fn=lambda r: fn(0))
    return server, make_clients(NumpyClippingClient, config, device, lambda: nn.Sequential(nn.Linear(features, 2)), customise)


# ---------------------------------------------------------------------------------------------------------------
# second stages and sub-variants of the representation-learning examples
# ---------------------------------------------------------------------------------------------------------------
def _flat_dim(config: dict[str, Any]) -> int:
    return 28 * 28 if config["dataset"] == "mnist" else 3 * 32 * 32


@scenario("fedprox_vae_example")
def fedprox_vae_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    """ae_examples/fedprox_vae_example: the VAE of ``ae_example`` trained with FedProx (adaptive proximal weight)."""
    from fl4health_b200.clients.fed_prox_client import FedProxClient
    from fl4health_b200.model_bases.autoencoders_base import VariationalAe
    from fl4health_b200.preprocessing.autoencoders.loss import VaeLoss
    from fl4health_b200.servers.adaptive_constraint_servers.fedprox_server import FedProxServer
    from fl4health_b200.utils.dataset_converter import AutoEncoderDatasetConverter

    in_dim, latent = _flat_dim(config), 16

    class Encoder(nn.Module):
        def __init__(self) -> None:
            super().__init__()
            self.body = nn.Sequential(nn.Flatten(), nn.Linear(in_dim, 64), nn.ReLU())
            self.mu, self.logvar = nn.Linear(64, latent), nn.Linear(64, latent)

        def forward(self, x: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
            h = self.body(x)
            return self.mu(h), self.logvar(h)

    class FlatVaeLoss(VaeLoss):
        def forward(self, preds: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
            return super().forward(preds, target.flatten(1))

    def customise(client: Any) -> None:
        def loaders(cfg: dict[str, Any]) -> tuple[Any, Any]:
            train, val = client_datasets(config, client.client_index)
            convert = lambda ds: AutoEncoderDatasetConverter().convert_dataset(ds)  # noqa: E731
            return BatchedTensorLoader(convert(train), config["batch_size"], shuffle=True), BatchedTensorLoader(convert(val), config["batch_size"])

        client.get_data_loaders = loaders
        client.get_criterion = lambda cfg: FlatVaeLoss(latent, nn.MSELoss(reduction="sum"))

    factory = lambda: VariationalAe(Encoder(), nn.Sequential(nn.Linear(latent, 64), nn.ReLU(), nn.Linear(64, in_dim)))  # noqa: E731
    clients = make_clients(FedProxClient, config, device, factory, customise, metrics=[])
    return _fl_server(config, _adaptive_strategy(config), FedProxServer), clients


def _pretrain_and_save(model: nn.Module, path: Path) -> None:
    path.parent.mkdir(parents=True, exist_ok=True)
    torch.save(model, path)


class _CondEncoder(nn.Module):
    """Module-level (picklable: the dimensionality-reduction processors ``torch.load`` whole models) CVAE encoder."""

    def __init__(self, in_dim: int, n_conditions: int, latent: int) -> None:
        super().__init__()
        self.body = nn.Sequential(nn.Linear(in_dim + n_conditions, 32), nn.ReLU())
        self.mu, self.logvar = nn.Linear(32, latent), nn.Linear(32, latent)

    def forward(self, x: torch.Tensor, condition: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        h = self.body(torch.cat((x.flatten(1), condition), dim=-1))
        return self.mu(h), self.logvar(h)


class _CondDecoder(nn.Module):
    def __init__(self, in_dim: int, n_conditions: int, latent: int) -> None:
        super().__init__()
        self.body = nn.Linear(latent + n_conditions, in_dim)

    def forward(self, z: torch.Tensor, condition: torch.Tensor) -> torch.Tensor:
        return self.body(torch.cat((z, condition), dim=-1))


@scenario("cvae_dim_example")
def cvae_dim_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    """ae_examples/cvae_dim_example: a (here: freshly initialised and saved) CVAE encoder reduces every sample to its
    latent code under a fixed client condition (``CvaeFixedConditionProcessor``); a small classifier is federated on
    the codes."""
    from fl4health_b200.clients.basic_client import BasicClient
    from fl4health_b200.model_bases.autoencoders_base import ConditionalVae
    from fl4health_b200.preprocessing.autoencoders.dim_reduction import CvaeFixedConditionProcessor
    from fl4health_b200.utils.dataset import TensorDataset

    in_dim, latent, n_conditions = _flat_dim(config), 8, int(config["n_clients"])
    out = Path(config.get("checkpoint_dir", "examples_out/cvae_dim"))

    torch.manual_seed(config["seed"])
    _pretrain_and_save(ConditionalVae(_CondEncoder(in_dim, n_conditions, latent), _CondDecoder(in_dim, n_conditions, latent)), out / "cvae.pt")

    def customise(client: Any) -> None:
        condition = torch.nn.functional.one_hot(torch.tensor(client.client_index), n_conditions).float()
        processor = CvaeFixedConditionProcessor(out / "cvae.pt", condition, device=torch.device("cpu"), return_mu_only=True)

        def reduce(ds: TensorDataset) -> TensorDataset:
            return TensorDataset(processor(ds.data.reshape(len(ds.data), -1).float()), ds.targets)

        def loaders(cfg: dict[str, Any]) -> tuple[Any, Any]:
            train, val = client_datasets(config, client.client_index)
            return BatchedTensorLoader(reduce(train), config["batch_size"], shuffle=True), BatchedTensorLoader(reduce(val), config["batch_size"])

        client.get_data_loaders = loaders

    factory = lambda: nn.Sequential(nn.Linear(latent, 32), nn.ReLU(), nn.Linear(32, 10))  # noqa: E731
    return _fl_server(config, BasicFedAvg(**strategy_kwargs(config))), make_clients(BasicClient, config, device, factory, customise)


@scenario("fedpca_dim_reduction_example")
def fedpca_dim_reduction_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    """fedpca_examples/dim_reduction: a saved ``PcaModule`` (what ``fedpca_example`` / perform_pca produces) projects the
    data onto its leading components (``PcaPreprocessor``); a classifier is federated on the projections."""
    from fl4health_b200.clients.basic_client import BasicClient
    from fl4health_b200.model_bases.pca import PcaModule
    from fl4health_b200.preprocessing.pca_preprocessor import PcaPreprocessor
    from fl4health_b200.utils.dataset import TensorDataset

    components = int(config.get("new_dimension", 16))
    out = Path(config.get("checkpoint_dir", "examples_out/fedpca_dim_reduction"))
    # stage 1 stand-in: principal components of one reference shard, saved the way FedPCAClient / the server do
    reference, _ = client_datasets(config, 0)
    pca = PcaModule(low_rank=False, full_svd=False, rank_estimation=components)
    principal_components, singular_values = pca(reference.data.reshape(len(reference.data), -1).float(), center_data=True)
    pca.set_principal_components(principal_components, singular_values)
    _pretrain_and_save(pca, out / "pca.pt")

    def customise(client: Any) -> None:
        def loaders(cfg: dict[str, Any]) -> tuple[Any, Any]:
            train, val = client_datasets(config, client.client_index)
            flat = lambda ds: TensorDataset(ds.data.reshape(len(ds.data), -1).float(), ds.targets)  # noqa: E731
            reducer = PcaPreprocessor(out / "pca.pt")
            return (BatchedTensorLoader(reducer.reduce_dimension(components, flat(train)), config["batch_size"], shuffle=True),
                    BatchedTensorLoader(reducer.reduce_dimension(components, flat(val)), config["batch_size"]))

        client.get_data_loaders = loaders

    factory = lambda: nn.Sequential(nn.Linear(components, 32), nn.ReLU(), nn.Linear(32, 10))  # noqa: E731
    return _fl_server(config, BasicFedAvg(**strategy_kwargs(config))), make_clients(BasicClient, config, device, factory, customise)


@scenario("fedsimclr_finetuning_example")
def fedsimclr_finetuning_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    """fedsimclr_example/fedsimclr_finetuning_example: the encoder saved by the pre-training stage is loaded with
    ``FedSimClrModel.load_pretrained_model`` (``pretrain=False``: encoder → prediction head) and fine-tuned with labels."""
    from examples.models import FeatureCnn
    from fl4health_b200.clients.basic_client import BasicClient
    from fl4health_b200.model_bases.fedsimclr_base import FedSimClrModel

    out = Path(config.get("checkpoint_dir", "examples_out/fedsimclr"))
    torch.manual_seed(config["seed"])
    pretrained = FedSimClrModel(FeatureCnn(config["dataset"]), nn.Linear(FeatureCnn.out_dim, 32), nn.Linear(FeatureCnn.out_dim, 10),
                                pretrain=True)  # stands in for the checkpoint written by `fedsimclr_example`
    _pretrain_and_save(pretrained, out / "pretrained.pt")
    factory = lambda: FedSimClrModel.load_pretrained_model(out / "pretrained.pt")  # noqa: E731
    return _fl_server(config, BasicFedAvg(**strategy_kwargs(config))), make_clients(BasicClient, config, device, factory)


@scenario("warmed_up_fenda_example")
def warmed_up_fenda_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    """warm_up_example/warmed_up_fenda: both feature extractors of a FENDA model start from a pretrained plain CNN; the
    weights mapping renames ``features.*`` to ``first_feature_extractor.*`` / ``second_feature_extractor.*``."""
    from examples.models import ConcatHead, FeatureCnn
    from fl4health_b200.clients.fenda_client import FendaClient
    from fl4health_b200.model_bases.fenda_base import FendaModel
    from fl4health_b200.preprocessing.warmed_up_module import WarmedUpModule

    import json

    out = Path(config.get("checkpoint_dir", "examples_out/warmed_up_fenda"))
    torch.manual_seed(config["seed"] + 1)
    pretrained = SmallCnn(config["dataset"])
    out.mkdir(parents=True, exist_ok=True)
    (out / "weights_mapping.json").write_text(json.dumps({"first_feature_extractor": "features", "second_feature_extractor": "features"}))

    def factory() -> nn.Module:
        model = FendaModel(FeatureCnn(config["dataset"]), FeatureCnn(config["dataset"]), ConcatHead())
        return WarmedUpModule(pretrained_model=pretrained, weights_mapping_path=out / "weights_mapping.json").load_from_pretrained(model)

    return _fl_server(config, BasicFedAvg(**strategy_kwargs(config))), make_clients(FendaClient, config, device, factory)
