"""Scenario registry: one builder per example of the reference's ``examples/`` tree.

Every builder receives the merged YAML config and a device and returns ``(server, clients)``; ``examples/run.py``
executes them.  Builders are intentionally short: the point of each example is which client class, server,
strategy and model wiring a method needs — data / optimizer / loss hooks come from ``ExampleClientMixin``.
"""

from __future__ import annotations

from collections.abc import Callable
from pathlib import Path
from typing import Any

import torch
from torch import nn

from examples.common import ExampleClientMixin, make_config_fn, strategy_kwargs
from examples.models import ClassifierHead, ConcatHead, FeatureCnn, SmallCnn
from fl4health_b200.common.typing import ndarrays_to_parameters
from fl4health_b200.metrics import Accuracy
from fl4health_b200.servers.base_server import FlServer
from fl4health_b200.servers.client_manager import SimpleClientManager
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg

Builder = Callable[[dict[str, Any], torch.device], tuple[Any, list[Any]]]
SCENARIOS: dict[str, Builder] = {}


def scenario(name: str) -> Callable[[Builder], Builder]:
    def register(fn: Builder) -> Builder:
        SCENARIOS[name] = fn
        return fn

    return register


def make_clients(client_cls: type, config: dict[str, Any], device: torch.device, model_factory: Callable[[], nn.Module],
                 customise: Callable[[Any], None] | None = None, metrics: list[Any] | None = None, **client_kwargs: Any) -> list[Any]:
    """K instances of ``class Example<Client>(ExampleClientMixin, client_cls)``."""
    if issubclass(client_cls, ExampleClientMixin):
        cls = client_cls
    else:
        cls = type(f"Example{client_cls.__name__}", (ExampleClientMixin, client_cls), {})
    clients = []
    for index in range(int(config["n_clients"])):
        client = cls(Path(config["data_dir"]), [Accuracy()] if metrics is None else metrics, device, client_name=f"client_{index}",
                     **client_kwargs)
        client.example_config, client.client_index, client.model_factory = config, index, model_factory
        if customise is not None:
            customise(client)
        clients.append(client)
    return clients


def _dict_optimizers(client: Any, parts: dict[str, Callable[[Any], Any]]) -> None:
    """Install a ``get_optimizer`` returning one optimizer per named sub-module."""
    client.get_optimizer = lambda config: {key: client.make_optimizer(select(client).parameters()) for key, select in parts.items()}


def _fl_server(config: dict[str, Any], strategy: Any, server_cls: type = FlServer, config_fn: Any = None, **kwargs: Any) -> Any:
    fn = config_fn or make_config_fn(config)
    return server_cls(SimpleClientManager(), {"n_server_rounds": config["n_server_rounds"]}, strategy,
                      on_init_parameters_config_fn=fn, **kwargs)


def _initial_parameters(model: nn.Module) -> Any:
    return ndarrays_to_parameters([v.detach().clone() for v in model.state_dict().values()])


# ---------------------------------------------------------------------------------------------------------------
# plain FedAvg family
# ---------------------------------------------------------------------------------------------------------------
@scenario("basic_example")
def basic_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.basic_client import BasicClient

    clients = make_clients(BasicClient, config, device, lambda: SmallCnn(config["dataset"]))
    return _fl_server(config, BasicFedAvg(**strategy_kwargs(config)), accept_failures=False), clients


@scenario("fedopt_example")
def fedopt_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.basic_client import BasicClient
    from fl4health_b200.strategies.fedopt import FedAdam

    template = SmallCnn(config["dataset"])
    torch.manual_seed(config["seed"])
    strategy = FedAdam(initial_parameters=_initial_parameters(SmallCnn(config["dataset"])), eta=config.get("server_learning_rate", 0.005),
                       **strategy_kwargs(config))
    del template
    return _fl_server(config, strategy), make_clients(BasicClient, config, device, lambda: SmallCnn(config["dataset"]))


@scenario("fedbn_example")
def fedbn_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.fedbn_client import FedBnClient
    from fl4health_b200.parameter_exchange.layer_exchanger import LayerExchangerWithExclusions

    def customise(client: Any) -> None:
        client.get_parameter_exchanger = lambda cfg: LayerExchangerWithExclusions(client.model, {nn.BatchNorm2d})

    clients = make_clients(FedBnClient, config, device, lambda: SmallCnn(config["dataset"], batch_norm=True), customise)
    return _fl_server(config, BasicFedAvg(**strategy_kwargs(config))), clients


@scenario("flash_example")
def flash_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.flash_client import FlashClient
    from fl4health_b200.strategies.flash import Flash

    config = {**config, "local_epochs": config.get("local_epochs", 2), "local_steps": None}
    fn = make_config_fn(config, gamma=config.get("gamma", 0.01))
    strategy = Flash(initial_parameters=None, eta=config.get("server_learning_rate", 0.005), **strategy_kwargs(config, fn))
    return _fl_server(config, strategy, config_fn=fn), make_clients(FlashClient, config, device, lambda: SmallCnn(config["dataset"]))


@scenario("ensemble_example")
def ensemble_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.ensemble_client import EnsembleClient
    from fl4health_b200.model_bases.ensemble_base import EnsembleModel

    def customise(client: Any) -> None:
        client.get_optimizer = lambda cfg: {k: client.make_optimizer(m.parameters()) for k, m in client.model.ensemble_models.items()}

    factory = lambda: EnsembleModel({f"model_{i}": SmallCnn(config["dataset"]) for i in range(3)})  # noqa: E731
    return _fl_server(config, BasicFedAvg(**strategy_kwargs(config))), make_clients(EnsembleClient, config, device, factory, customise)


# ---------------------------------------------------------------------------------------------------------------
# drift-constrained / variance-reduced methods
# ---------------------------------------------------------------------------------------------------------------
def _adaptive_strategy(config: dict[str, Any]) -> Any:
    from fl4health_b200.strategies.fedavg_with_adaptive_constraint import FedAvgWithAdaptiveConstraint

    return FedAvgWithAdaptiveConstraint(
        initial_parameters=None, initial_loss_weight=config.get("initial_loss_weight", 0.1),
        adapt_loss_weight=config.get("adapt_loss_weight", True), loss_weight_delta=config.get("loss_weight_delta", 0.05),
        loss_weight_patience=config.get("loss_weight_patience", 2), **strategy_kwargs(config))


@scenario("fedprox_example")
def fedprox_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.fed_prox_client import FedProxClient
    from fl4health_b200.servers.adaptive_constraint_servers.fedprox_server import FedProxServer

    clients = make_clients(FedProxClient, config, device, lambda: SmallCnn(config["dataset"]))
    return _fl_server(config, _adaptive_strategy(config), FedProxServer), clients


@scenario("scaffold_example")
def scaffold_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.scaffold_client import ScaffoldClient
    from fl4health_b200.servers.scaffold_server import ScaffoldServer
    from fl4health_b200.strategies.scaffold import Scaffold

    factory = lambda: SmallCnn(config["dataset"], batch_norm=True, frozen_conv=True)  # noqa: E731
    torch.manual_seed(config["seed"])
    template = factory()
    strategy = Scaffold(initial_parameters=_initial_parameters(template), model=template, learning_rate=1.0,
                        **{k: v for k, v in strategy_kwargs(config).items() if k not in ("min_fit_clients", "min_evaluate_clients")})
    server = ScaffoldServer(SimpleClientManager(), {"n_server_rounds": config["n_server_rounds"]}, strategy)
    return server, make_clients(ScaffoldClient, config, device, factory)


@scenario("ditto_example")
def ditto_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.ditto_client import DittoClient
    from fl4health_b200.servers.adaptive_constraint_servers.ditto_server import DittoServer

    def customise(client: Any) -> None:
        _dict_optimizers(client, {"global": lambda c: c.global_model, "local": lambda c: c.model})

    clients = make_clients(DittoClient, config, device, lambda: SmallCnn(config["dataset"]), customise)
    return _fl_server(config, _adaptive_strategy({**config, "adapt_loss_weight": False}), DittoServer), clients


@scenario("ditto_example_dynamic")
def ditto_example_dynamic(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    """Ditto obtained by ``make_it_personal(FlexibleClient, DITTO)`` instead of the dedicated client class."""
    from fl4health_b200.clients.flexible import FlexibleClient
    from fl4health_b200.mixins import PersonalizedMode, make_it_personal
    from fl4health_b200.servers.adaptive_constraint_servers.ditto_server import DittoServer

    client_cls = make_it_personal(type("ExampleFlexible", (ExampleClientMixin, FlexibleClient), {}), PersonalizedMode.DITTO)
    clients = make_clients(client_cls, config, device, lambda: SmallCnn(config["dataset"]))
    return _fl_server(config, _adaptive_strategy({**config, "adapt_loss_weight": False}), DittoServer), clients


@scenario("mr_mtl_example")
def mr_mtl_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.mr_mtl_client import MrMtlClient
    from fl4health_b200.servers.adaptive_constraint_servers.mrmtl_server import MrMtlServer

    clients = make_clients(MrMtlClient, config, device, lambda: SmallCnn(config["dataset"]))
    return _fl_server(config, _adaptive_strategy({**config, "adapt_loss_weight": False}), MrMtlServer), clients


@scenario("ditto_mkmmd_example")
def ditto_mkmmd_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.mkmmd_clients import DittoMkMmdClient
    from fl4health_b200.servers.adaptive_constraint_servers.ditto_server import DittoServer

    def customise(client: Any) -> None:
        _dict_optimizers(client, {"global": lambda c: c.global_model, "local": lambda c: c.model})

    clients = make_clients(DittoMkMmdClient, config, device, lambda: SmallCnn(config["dataset"]), customise,
                           mkmmd_loss_weight=1.0, feature_extraction_layers=["features"], beta_global_update_interval=2,
                           num_accumulating_batches=2)
    return _fl_server(config, _adaptive_strategy({**config, "adapt_loss_weight": False}), DittoServer), clients


# ---------------------------------------------------------------------------------------------------------------
# personalised architectures
# ---------------------------------------------------------------------------------------------------------------
@scenario("apfl_example")
def apfl_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.apfl_client import ApflClient
    from fl4health_b200.model_bases.apfl_base import ApflModule

    def customise(client: Any) -> None:
        _dict_optimizers(client, {"local": lambda c: c.model.local_model, "global": lambda c: c.model.global_model})

    clients = make_clients(ApflClient, config, device, lambda: ApflModule(SmallCnn(config["dataset"]), alpha_lr=0.1), customise)
    return _fl_server(config, BasicFedAvg(**strategy_kwargs(config))), clients


@scenario("feddg_ga_example")
def feddg_ga_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.client_managers.fixed_sampling_client_manager import FixedSamplingClientManager
    from fl4health_b200.clients.apfl_client import ApflClient
    from fl4health_b200.model_bases.apfl_base import ApflModule
    from fl4health_b200.strategies.feddg_ga import FedDgGa

    def customise(client: Any) -> None:
        _dict_optimizers(client, {"local": lambda c: c.model.local_model, "global": lambda c: c.model.global_model})

    fn = make_config_fn(config, evaluate_after_fit=True, pack_losses_with_val_metrics=True)
    kwargs = {k: v for k, v in strategy_kwargs(config, fn).items() if not k.startswith("min_")}
    server = FlServer(FixedSamplingClientManager(), {"n_server_rounds": config["n_server_rounds"]}, FedDgGa(**kwargs),
                      on_init_parameters_config_fn=fn)
    return server, make_clients(ApflClient, config, device, lambda: ApflModule(SmallCnn(config["dataset"])), customise)


def _split_model(dataset: str, cls: type) -> nn.Module:
    return cls(FeatureCnn(dataset), ClassifierHead())


@scenario("fedper_example")
def fedper_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.fedper_client import FedPerClient
    from fl4health_b200.model_bases.sequential_split_models import SequentiallySplitExchangeBaseModel

    clients = make_clients(FedPerClient, config, device, lambda: _split_model(config["dataset"], SequentiallySplitExchangeBaseModel))
    return _fl_server(config, BasicFedAvg(**strategy_kwargs(config))), clients


@scenario("fedrep_example")
def fedrep_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.fedrep_client import FedRepClient
    from fl4health_b200.model_bases.fedrep_base import FedRepModel

    def customise(client: Any) -> None:
        _dict_optimizers(client, {"representation": lambda c: c.model.base_module, "head": lambda c: c.model.head_module})

    def fn(server_round: int) -> dict[str, Any]:
        return {"current_server_round": server_round, "local_head_steps": 2, "local_rep_steps": 2, "batch_size": config["batch_size"]}

    clients = make_clients(FedRepClient, config, device, lambda: _split_model(config["dataset"], FedRepModel), customise)
    return _fl_server(config, BasicFedAvg(**strategy_kwargs(config, fn)), config_fn=fn), clients


@scenario("moon_example")
def moon_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.moon_client import MoonClient
    from fl4health_b200.model_bases.moon_base import MoonModel

    clients = make_clients(MoonClient, config, device, lambda: _split_model(config["dataset"], MoonModel))
    return _fl_server(config, BasicFedAvg(**strategy_kwargs(config))), clients


def _parallel(dataset: str, cls: type, **kwargs: Any) -> nn.Module:
    return cls(FeatureCnn(dataset), FeatureCnn(dataset), ConcatHead(), **kwargs)


@scenario("fenda_example")
def fenda_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.fenda_client import FendaClient
    from fl4health_b200.model_bases.fenda_base import FendaModel

    clients = make_clients(FendaClient, config, device, lambda: _parallel(config["dataset"], FendaModel))
    return _fl_server(config, BasicFedAvg(**strategy_kwargs(config))), clients


@scenario("perfcl_example")
def perfcl_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.perfcl_client import PerFclClient
    from fl4health_b200.model_bases.perfcl_base import PerFclModel

    clients = make_clients(PerFclClient, config, device, lambda: _parallel(config["dataset"], PerFclModel))
    return _fl_server(config, BasicFedAvg(**strategy_kwargs(config))), clients


@scenario("fenda_ditto_example")
def fenda_ditto_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.fenda_ditto_client import FendaDittoClient
    from fl4health_b200.model_bases.fenda_base import FendaModel
    from fl4health_b200.model_bases.sequential_split_models import SequentiallySplitModel
    from fl4health_b200.servers.adaptive_constraint_servers.ditto_server import DittoServer

    def customise(client: Any) -> None:
        client.get_global_model = lambda cfg: SequentiallySplitModel(FeatureCnn(config["dataset"]), ClassifierHead())
        _dict_optimizers(client, {"global": lambda c: c.global_model, "local": lambda c: c.model})

    clients = make_clients(FendaDittoClient, config, device, lambda: _parallel(config["dataset"], FendaModel), customise)
    return _fl_server(config, _adaptive_strategy({**config, "adapt_loss_weight": False}), DittoServer), clients


@scenario("gpfl_example")
def gpfl_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.gpfl_client import GpflClient
    from fl4health_b200.model_bases.gpfl_base import GpflModel

    def customise(client: Any) -> None:
        _dict_optimizers(client, {"model": lambda c: c.model.gpfl_main_module, "gce": lambda c: c.model.gce, "cov": lambda c: c.model.cov})

    factory = lambda: GpflModel(FeatureCnn(config["dataset"]), ClassifierHead(), feature_dim=FeatureCnn.out_dim, num_classes=10)  # noqa: E731
    return _fl_server(config, BasicFedAvg(**strategy_kwargs(config))), make_clients(GpflClient, config, device, factory, customise)


@scenario("fedpm_example")
def fedpm_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.fedpm_client import FedPmClient
    from fl4health_b200.servers.fedpm_server import FedPmServer
    from fl4health_b200.strategies.fedpm import FedPm

    fn = make_config_fn(config, is_masked_model=False)
    server = FedPmServer(SimpleClientManager(), {"n_server_rounds": config["n_server_rounds"]}, FedPm(**strategy_kwargs(config, fn)),
                         reset_frequency=config.get("priors_reset_frequency", 2), on_init_parameters_config_fn=fn)
    clients = make_clients(FedPmClient, {**config, "learning_rate": 0.5}, device, lambda: SmallCnn(config["dataset"]))
    return server, clients


# ---------------------------------------------------------------------------------------------------------------
# partial exchange
# ---------------------------------------------------------------------------------------------------------------
@scenario("dynamic_layer_exchange_example")
def dynamic_layer_exchange_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.partial_weight_exchange_client import PartialWeightExchangeClient
    from fl4health_b200.parameter_exchange.layer_exchanger import DynamicLayerExchanger
    from fl4health_b200.parameter_exchange.parameter_selection_criteria import LayerSelectionFunctionConstructor
    from fl4health_b200.strategies.fedavg_dynamic_layer import FedAvgDynamicLayer

    def customise(client: Any) -> None:
        client.store_initial_model = True
        selector = LayerSelectionFunctionConstructor(1e-9, config.get("exchange_percentage", 0.5), normalize=False).select_by_percentage()
        client.get_parameter_exchanger = lambda cfg: DynamicLayerExchanger(selector)

    clients = make_clients(PartialWeightExchangeClient, config, device, lambda: SmallCnn(config["dataset"]), customise)
    return _fl_server(config, FedAvgDynamicLayer(**strategy_kwargs(config))), clients


@scenario("sparse_tensor_partial_exchange_example")
def sparse_tensor_partial_exchange_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.partial_weight_exchange_client import PartialWeightExchangeClient
    from fl4health_b200.parameter_exchange.parameter_selection_criteria import largest_final_magnitude_scores
    from fl4health_b200.parameter_exchange.sparse_coo_parameter_exchanger import SparseCooParameterExchanger
    from fl4health_b200.strategies.fedavg_sparse_coo_tensor import FedAvgSparseCooTensor

    def customise(client: Any) -> None:
        client.store_initial_model = True
        client.get_parameter_exchanger = lambda cfg: SparseCooParameterExchanger(config.get("sparsity_level", 0.3), largest_final_magnitude_scores)

    clients = make_clients(PartialWeightExchangeClient, config, device, lambda: SmallCnn(config["dataset"]), customise)
    return _fl_server(config, FedAvgSparseCooTensor(**strategy_kwargs(config))), clients


# ---------------------------------------------------------------------------------------------------------------
# differential privacy
# ---------------------------------------------------------------------------------------------------------------
def _dp_fn(config: dict[str, Any]) -> Callable[[int], dict[str, Any]]:
    return make_config_fn(config, clipping_bound=config.get("clipping_bound", 1.0), noise_multiplier=config.get("noise_multiplier", 0.5),
                          adaptive_clipping=config.get("adaptive_clipping", True))


@scenario("instance_level_dp_example")
def instance_level_dp_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.client_managers.poisson_sampling_manager import PoissonSamplingClientManager
    from fl4health_b200.clients.instance_level_dp_client import InstanceLevelDpClient
    from fl4health_b200.servers.instance_level_dp_server import InstanceLevelDpServer

    fn = _dp_fn(config)
    kwargs = {**strategy_kwargs(config, fn), "fraction_fit": 1.0, "fraction_evaluate": 1.0}
    server = InstanceLevelDpServer(
        PoissonSamplingClientManager(), {"n_server_rounds": config["n_server_rounds"]}, noise_multiplier=config.get("noise_multiplier", 0.5),
        batch_size=config["batch_size"], num_server_rounds=config["n_server_rounds"], strategy=BasicFedAvg(**kwargs),
        local_steps=config["local_steps"], on_init_parameters_config_fn=fn)
    return server, make_clients(InstanceLevelDpClient, config, device, lambda: SmallCnn(config["dataset"], batch_norm=True))


@scenario("client_level_dp_example")
def client_level_dp_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.client_managers.poisson_sampling_manager import PoissonSamplingClientManager
    from fl4health_b200.clients.clipping_client import NumpyClippingClient
    from fl4health_b200.servers.client_level_dp_fed_avg_server import ClientLevelDPFedAvgServer
    from fl4health_b200.strategies.client_dp_fedavgm import ClientLevelDPFedAvgM

    fn = _dp_fn(config)
    kwargs = {k: v for k, v in strategy_kwargs(config, fn).items() if k not in ("min_fit_clients", "min_evaluate_clients")}
    strategy = ClientLevelDPFedAvgM(fraction_fit=1.0, fraction_evaluate=1.0, adaptive_clipping=True, initial_clipping_bound=0.5,
                                    weight_noise_multiplier=0.1, clipping_noise_multiplier=5.0, weighted_aggregation=True, **kwargs)
    server = ClientLevelDPFedAvgServer(PoissonSamplingClientManager(), {"n_server_rounds": config["n_server_rounds"]}, strategy,
                                       server_noise_multiplier=0.1, num_server_rounds=config["n_server_rounds"],
                                       on_init_parameters_config_fn=lambda r: fn(0))
    return server, make_clients(NumpyClippingClient, config, device, lambda: SmallCnn(config["dataset"]))


# ---------------------------------------------------------------------------------------------------------------
# representation learning / preprocessing
# ---------------------------------------------------------------------------------------------------------------
@scenario("fedpca_example")
def fedpca_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from examples.common import client_datasets
    from fl4health_b200.clients.fed_pca_client import FedPCAClient
    from fl4health_b200.engine.data import BatchedTensorLoader
    from fl4health_b200.strategies.fedpca import FedPCA
    from fl4health_b200.utils.dataset import TensorDataset

    out_dir = Path(config.get("output_dir", "examples/outputs"))
    out_dir.mkdir(parents=True, exist_ok=True)

    class Client(FedPCAClient):
        def get_data_loaders(self, cfg: dict[str, Any]) -> tuple[Any, Any]:
            train, val = client_datasets(config, self.client_index)
            flat = lambda ds: TensorDataset(ds.data.reshape(len(ds.data), -1).float(), ds.targets)  # noqa: E731
            return BatchedTensorLoader(flat(train), config["batch_size"]), BatchedTensorLoader(flat(val), config["batch_size"])

    def fn(server_round: int) -> dict[str, Any]:
        return {"current_server_round": server_round, "low_rank": True, "full_svd": False, "rank_estimation": 8, "center_data": True,
                "num_components_eval": 8}

    clients = []
    for index in range(int(config["n_clients"])):
        client = Client(Path(config["data_dir"]), device, out_dir, client_name=f"client_{index}")
        client.client_index = index
        clients.append(client)
    n = int(config["n_clients"])
    strategy = FedPCA(min_fit_clients=n, min_evaluate_clients=n, min_available_clients=n, on_fit_config_fn=fn, on_evaluate_config_fn=fn)
    return FlServer(SimpleClientManager(), {"n_server_rounds": 1}, strategy, on_init_parameters_config_fn=fn), clients


@scenario("ae_example")
def ae_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    """Federated VAE: ``VariationalAe`` + ``VaeLoss`` + ``AutoEncoderDatasetConverter`` (target := input)."""
    from examples.common import client_datasets
    from fl4health_b200.clients.basic_client import BasicClient
    from fl4health_b200.engine.data import BatchedTensorLoader
    from fl4health_b200.model_bases.autoencoders_base import VariationalAe
    from fl4health_b200.preprocessing.autoencoders.loss import VaeLoss
    from fl4health_b200.utils.dataset_converter import AutoEncoderDatasetConverter

    in_dim, latent = 28 * 28 if config["dataset"] == "mnist" else 3 * 32 * 32, 16

    class Encoder(nn.Module):
        def __init__(self) -> None:
            super().__init__()
            self.body = nn.Sequential(nn.Flatten(), nn.Linear(in_dim, 64), nn.ReLU())
            self.mu, self.logvar = nn.Linear(64, latent), nn.Linear(64, latent)

        def forward(self, x: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
            h = self.body(x)
            return self.mu(h), self.logvar(h)

    decoder = lambda: nn.Sequential(nn.Linear(latent, 64), nn.ReLU(), nn.Linear(64, in_dim))  # noqa: E731

    def customise(client: Any) -> None:
        def loaders(cfg: dict[str, Any]) -> tuple[Any, Any]:
            train, val = client_datasets(config, client.client_index)
            convert = lambda ds: AutoEncoderDatasetConverter().convert_dataset(ds)  # noqa: E731
            return BatchedTensorLoader(convert(train), config["batch_size"], shuffle=True), BatchedTensorLoader(convert(val), config["batch_size"])

        client.get_data_loaders = loaders
        client.get_criterion = lambda cfg: VaeLoss(latent, nn.MSELoss(reduction="sum"))

    clients = make_clients(BasicClient, config, device, lambda: VariationalAe(Encoder(), decoder()), customise, metrics=[])
    return _fl_server(config, BasicFedAvg(**strategy_kwargs(config))), clients


@scenario("warm_up_example")
def warm_up_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    """FedProx clients whose models start from a (differently named) pretrained network via ``WarmedUpModule``."""
    from fl4health_b200.clients.fed_prox_client import FedProxClient
    from fl4health_b200.preprocessing.warmed_up_module import WarmedUpModule
    from fl4health_b200.servers.adaptive_constraint_servers.fedprox_server import FedProxServer

    torch.manual_seed(config["seed"] + 1)
    pretrained = SmallCnn(config["dataset"])

    def factory() -> nn.Module:
        return WarmedUpModule(pretrained_model=pretrained).load_from_pretrained(SmallCnn(config["dataset"]))

    return _fl_server(config, _adaptive_strategy(config), FedProxServer), make_clients(FedProxClient, config, device, factory)


# ---------------------------------------------------------------------------------------------------------------
# more scenarios: text, evaluation-only, one-shot merging, fine-tuning, SSL, tabular alignment, segmentation
# ---------------------------------------------------------------------------------------------------------------
@scenario("bert_finetuning_example")
def bert_finetuning_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    """Sequence classification with a BERT encoder and dict inputs (``input_ids`` / ``attention_mask``).  The default
    config uses ``BertConfig.tiny`` on synthetic token sequences; set ``bert_size: base`` for bert-base-cased shapes."""
    from torch.utils.data import DataLoader

    from fl4health_b200.clients.basic_client import BasicClient
    from fl4health_b200.models.bert import BertConfig, BertForSequenceClassification
    from fl4health_b200.utils.dataset import DictionaryDataset

    vocab, seq_len, classes = int(config.get("vocab_size", 1000)), int(config.get("seq_len", 32)), 4
    bert_cfg = BertConfig() if config.get("bert_size") == "base" else BertConfig.tiny(vocab)

    def dataset(n: int, seed: int) -> DictionaryDataset:
        gen = torch.Generator().manual_seed(seed)
        labels = torch.randint(0, classes, (n,), generator=gen)
        ids = torch.randint(10, vocab, (n, seq_len), generator=gen)
        ids[:, 1] = labels + 1  # a label-revealing token: makes the synthetic task learnable
        lengths = torch.randint(seq_len // 2, seq_len + 1, (n,), generator=gen)
        mask = (torch.arange(seq_len)[None, :] < lengths[:, None]).long()
        return DictionaryDataset({"input_ids": list(ids), "attention_mask": list(mask)}, labels)

    def customise(client: Any) -> None:
        def loaders(cfg: dict[str, Any]) -> tuple[Any, Any]:
            seed = config["seed"] + client.client_index
            return (DataLoader(dataset(int(config["samples_per_client"]), seed), batch_size=config["batch_size"], shuffle=True),
                    DataLoader(dataset(int(config["val_samples_per_client"]), 10_000 + seed), batch_size=config["batch_size"]))

        client.get_data_loaders = loaders

    factory = lambda: BertForSequenceClassification(bert_cfg, classes)  # noqa: E731
    clients = make_clients(BasicClient, {**config, "optimizer": "adamw", "learning_rate": config.get("learning_rate", 1e-3)}, device,
                           factory, customise)
    return _fl_server(config, BasicFedAvg(**strategy_kwargs(config))), clients


@scenario("dp_scaffold_example")
def dp_scaffold_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.client_managers.poisson_sampling_manager import PoissonSamplingClientManager
    from fl4health_b200.clients.scaffold_client import DPScaffoldClient
    from fl4health_b200.privacy.dp_engine import GradSampleModule, ModuleValidator
    from fl4health_b200.servers.scaffold_server import DPScaffoldServer
    from fl4health_b200.strategies.scaffold import OpacusScaffold

    fn = _dp_fn(config)
    factory = lambda: SmallCnn(config["dataset"], batch_norm=True)  # noqa: E731
    torch.manual_seed(config["seed"])
    template = GradSampleModule(ModuleValidator.fix(factory()))
    kwargs = {k: v for k, v in strategy_kwargs(config, fn).items() if k not in ("min_fit_clients", "min_evaluate_clients")}
    strategy = OpacusScaffold(model=template, fraction_fit=1.0, fraction_evaluate=1.0, learning_rate=1.0, **kwargs)
    server = DPScaffoldServer(PoissonSamplingClientManager(), {"n_server_rounds": config["n_server_rounds"]},
                              noise_multiplier=config.get("noise_multiplier", 0.5), batch_size=config["batch_size"],
                              num_server_rounds=config["n_server_rounds"], strategy=strategy, local_steps=config["local_steps"])
    return server, make_clients(DPScaffoldClient, config, device, factory)


@scenario("fl_plus_local_ft_example")
def fl_plus_local_ft_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    """FedAvg, then each client fine-tunes the final global model locally for a few steps at shutdown time."""
    from fl4health_b200.clients.basic_client import BasicClient

    class FineTuningClient(ExampleClientMixin, BasicClient):
        def shutdown(self) -> None:
            if getattr(self, "initialized", False):
                self.train_by_steps(int(self.example_config.get("local_ft_steps", 4)), current_round=None)
                loss, metrics = self.validate()
                self.fine_tuned_result = (loss, metrics)
            super().shutdown()

    clients = make_clients(FineTuningClient, config, device, lambda: SmallCnn(config["dataset"]))
    return _fl_server(config, BasicFedAvg(**strategy_kwargs(config))), clients


@scenario("fedsimclr_example")
def fedsimclr_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    """Federated SimCLR pre-training: ``SslTensorDataset`` makes (view, view') pairs, ``NtXentLoss`` is the criterion."""
    from examples.common import client_datasets
    from fl4health_b200.clients.basic_client import BasicClient
    from fl4health_b200.engine.data import BatchedTensorLoader
    from fl4health_b200.losses.contrastive_loss import NtXentLoss
    from fl4health_b200.model_bases.fedsimclr_base import FedSimClrModel
    from fl4health_b200.utils.dataset import SslTensorDataset

    def augment(x: torch.Tensor) -> torch.Tensor:
        return x + 0.1 * torch.randn_like(x)

    def customise(client: Any) -> None:
        def loaders(cfg: dict[str, Any]) -> tuple[Any, Any]:
            train, val = client_datasets(config, client.client_index)
            ssl = lambda ds: SslTensorDataset(ds.data, None, transform=None, target_transform=augment)  # noqa: E731
            return BatchedTensorLoader(ssl(train), config["batch_size"], shuffle=True), BatchedTensorLoader(ssl(val), config["batch_size"])

        client.get_data_loaders = loaders
        client.get_criterion = lambda cfg: NtXentLoss(client.device)
        # the "target" is the second view: embed it with the same model before the loss
        client.transform_target = lambda target: client.model(target)

    factory = lambda: FedSimClrModel(FeatureCnn(config["dataset"]), nn.Linear(FeatureCnn.out_dim, 32), pretrain=True)  # noqa: E731
    clients = make_clients(BasicClient, config, device, factory, customise, metrics=[])
    return _fl_server(config, BasicFedAvg(**strategy_kwargs(config))), clients


@scenario("feature_alignment_example")
def feature_alignment_example(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    """Tabular clients with heterogeneous columns aligned through ``TabularFeatureAlignmentServer`` (synthetic EHR-like
    frames; the reference uses a MIMIC-III sample)."""
    import numpy as np
    import pandas as pd

    from fl4health_b200.clients.tabular_data_client import TabularDataClient
    from fl4health_b200.engine.data import BatchedTensorLoader
    from fl4health_b200.servers.tabular_feature_alignment_server import TabularFeatureAlignmentServer
    from fl4health_b200.utils.dataset import TensorDataset

    def frame(seed: int, drop: str | None) -> pd.DataFrame:
        rng = np.random.default_rng(seed)
        n = int(config["samples_per_client"])
        age, smoker = rng.normal(60, 12, n), rng.integers(0, 2, n)
        df = pd.DataFrame({"patient_id": np.arange(n) + 10_000 * seed, "age": age, "smoker": smoker,
                           "admission": rng.choice(["emergency", "elective", "urgent"], n),
                           "mortality": ((age > 62) ^ (smoker == 1)).astype(int)})
        return df.drop(columns=[drop]) if drop else df

    class Client(TabularDataClient):
        def get_data_frame(self, cfg: dict[str, Any]) -> pd.DataFrame:
            return frame(config["seed"] + self.client_index, "admission" if self.client_index % 2 else None)

        def get_data_loaders(self, cfg: dict[str, Any]) -> tuple[Any, Any]:
            x = torch.from_numpy(np.asarray(self.aligned_features, dtype=np.float32))
            y = torch.from_numpy(np.asarray(self.aligned_targets)).long().reshape(-1)
            split = int(0.8 * len(x))
            return (BatchedTensorLoader(TensorDataset(x[:split], y[:split]), config["batch_size"], shuffle=True),
                    BatchedTensorLoader(TensorDataset(x[split:], y[split:]), config["batch_size"]))

        def get_model(self, cfg: dict[str, Any]) -> nn.Module:
            return nn.Sequential(nn.Linear(self.input_dimension, 32), nn.ReLU(), nn.Linear(32, self.output_dimension))

        def get_criterion(self, cfg: dict[str, Any]) -> nn.Module:
            return nn.CrossEntropyLoss()

        def get_optimizer(self, cfg: dict[str, Any]) -> torch.optim.Optimizer:
            return torch.optim.SGD(self.model.parameters(), lr=config["learning_rate"])

    clients = []
    for index in range(int(config["n_clients"])):
        client = Client(Path(config["data_dir"]), [Accuracy()], device, id_column="patient_id", targets="mortality", client_name=f"client_{index}")
        client.client_index = index
        clients.append(client)

    def initialize_parameters(input_dim: int, output_dim: int) -> Any:
        torch.manual_seed(config["seed"])
        return _initial_parameters(nn.Sequential(nn.Linear(input_dim, 32), nn.ReLU(), nn.Linear(32, output_dim)))

    kwargs = {k: v for k, v in strategy_kwargs(config).items() if k not in ("on_fit_config_fn", "on_evaluate_config_fn")}
    fl_config = {"n_server_rounds": config["n_server_rounds"], "batch_size": config["batch_size"], "local_steps": config["local_steps"]}
    return TabularFeatureAlignmentServer(SimpleClientManager(), fl_config, initialize_parameters, BasicFedAvg(**kwargs)), clients


@scenario("mr_mtl_example_flexible")
def mr_mtl_example_flexible(config: dict[str, Any], device: torch.device) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.flexible import FlexibleClient
    from fl4health_b200.mixins import PersonalizedMode, make_it_personal
    from fl4health_b200.servers.adaptive_constraint_servers.mrmtl_server import MrMtlServer

    client_cls = make_it_personal(type("ExampleFlexible", (ExampleClientMixin, FlexibleClient), {}), PersonalizedMode.MR_MTL)
    clients = make_clients(client_cls, config, device, lambda: SmallCnn(config["dataset"]))
    return _fl_server(config, _adaptive_strategy({**config, "adapt_loss_weight": False}), MrMtlServer), clients
