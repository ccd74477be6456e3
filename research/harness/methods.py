"""Method registry: ``METHODS[name](ctx) -> (server, clients)``.

A method is the FL4Health triple (client class, server class, strategy) plus the model family it needs; the task
supplies the building blocks (``features`` / ``head`` / ``parallel_head``).  This is the one place that replaces the
per-method ``server.py`` / ``client.py`` pairs of the reference's ``research/{cifar10,flamby,rxrx1,synthetic_data}``.

``lam`` is the method's penalty weight (FedProx / Ditto / MR-MTL λ, MOON / PerFCL μ); ``mmd_weight`` the MK-MMD /
Deep-MMD weight; both come from the ``ExperimentSpec`` and are what the sweeps vary next to the learning rate.
"""

from __future__ import annotations

from collections.abc import Callable
from dataclasses import dataclass
from pathlib import Path
from typing import Any

import torch
from torch import nn

from fl4health_b200.checkpointing.checkpointer import BestLossTorchModuleCheckpointer, LatestTorchModuleCheckpointer
from fl4health_b200.checkpointing.client_module import ClientCheckpointAndStateModule
from fl4health_b200.servers.client_manager import SimpleClientManager
from fl4health_b200.common.typing import Config, ndarrays_to_parameters
from fl4health_b200.engine.data import BatchedTensorLoader
from fl4health_b200.metrics.metric_aggregation import evaluate_metrics_aggregation_fn, fit_metrics_aggregation_fn
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg
from fl4health_b200.utils.dataset import TensorDataset
from research.harness.servers import FullExchangeServer, PersonalServer, make_personal
from research.harness.tasks import Task


@dataclass
class MethodContext:
    spec: Any  # ExperimentSpec
    task: Task
    device: torch.device
    run_dir: Path


class ResearchClientMixin:
    """Data / model / optimizer hooks driven by the experiment spec (combined as ``class C(ResearchClientMixin, X)``)."""

    ctx: MethodContext
    client_index: int = 0
    model_factory: Callable[[], nn.Module]

    def _loader(self, dataset: TensorDataset, shuffle: bool) -> BatchedTensorLoader:
        placement = "device" if self.device.type == "cuda" else "host"  # type: ignore[attr-defined]
        gen = torch.Generator().manual_seed(self.ctx.spec.seed * 100 + self.client_index) if shuffle else None
        return BatchedTensorLoader(dataset, self.ctx.spec.batch_size, shuffle=shuffle, placement=placement, device=self.device,  # type: ignore[attr-defined]
                                   generator=gen)

    def client_triple(self) -> tuple[TensorDataset, TensorDataset, TensorDataset]:
        return self.ctx.task.client_data(self.client_index, self.ctx.spec)

    def get_model(self, config: Config) -> nn.Module:
        torch.manual_seed(self.ctx.spec.seed)  # same initial weights everywhere, as a server broadcast would give
        return self.model_factory()

    def get_data_loaders(self, config: Config) -> tuple[BatchedTensorLoader, BatchedTensorLoader]:
        train, val, _ = self.client_triple()
        return self._loader(train, True), self._loader(val, False)

    def get_test_data_loader(self, config: Config) -> BatchedTensorLoader | None:
        return self._loader(self.client_triple()[2], False) if self.ctx.spec.evaluate_test_each_round else None

    def get_criterion(self, config: Config) -> nn.Module:
        return self.ctx.task.criterion()

    def make_optimizer(self, params: Any) -> torch.optim.Optimizer:
        spec = self.ctx.spec
        if spec.optimizer == "adamw":
            return torch.optim.AdamW(params, lr=spec.lr)
        return torch.optim.SGD(params, lr=spec.lr, momentum=spec.momentum, weight_decay=spec.weight_decay)

    def get_optimizer(self, config: Config) -> Any:
        return self.make_optimizer(self.model.parameters())  # type: ignore[attr-defined]


class PooledDataMixin(ResearchClientMixin):
    """``central`` baseline: one participant holding the union of every client's data."""

    def client_triple(self) -> tuple[TensorDataset, TensorDataset, TensorDataset]:
        triples = [self.ctx.task.client_data(i, self.ctx.spec) for i in range(self.ctx.task.n_clients)]
        join = lambda k: TensorDataset(torch.cat([t[k].data for t in triples]), torch.cat([t[k].targets for t in triples]))  # noqa: E731
        return join(0), join(1), join(2)


METHODS: dict[str, Callable[[MethodContext], tuple[Any, list[Any]]]] = {}


def method(name: str) -> Callable[[Callable[[MethodContext], tuple[Any, list[Any]]]], Callable[[MethodContext], tuple[Any, list[Any]]]]:
    def register(builder: Callable[[MethodContext], tuple[Any, list[Any]]]) -> Callable[[MethodContext], tuple[Any, list[Any]]]:
        METHODS[name] = builder
        return builder

    return register


# --------------------------------------------------------------------------------------------------------------- helpers
def config_fn(ctx: MethodContext, **extra: Any) -> Callable[[int], Config]:
    spec = ctx.spec

    def fn(server_round: int) -> Config:
        out: Config = {"current_server_round": server_round, "batch_size": spec.batch_size, "n_server_rounds": spec.rounds, **extra}
        out["local_epochs" if spec.local_epochs else "local_steps"] = spec.local_epochs or spec.local_steps
        return out

    return fn


def strategy_kwargs(ctx: MethodContext, fn: Callable[[int], Config] | None = None, n_clients: int | None = None) -> dict[str, Any]:
    n = n_clients or ctx.task.n_clients
    fn = fn or config_fn(ctx)
    return dict(min_fit_clients=n, min_evaluate_clients=n, min_available_clients=n, on_fit_config_fn=fn, on_evaluate_config_fn=fn,
                fit_metrics_aggregation_fn=fit_metrics_aggregation_fn, evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn)


def make_clients(ctx: MethodContext, client_cls: type, model_factory: Callable[[], nn.Module], customise: Callable[[Any], None] | None = None,
                 n_clients: int | None = None, mixin: type = ResearchClientMixin, **client_kwargs: Any) -> list[Any]:
    cls = type(f"Research{client_cls.__name__}", (mixin, client_cls), {})
    clients = []
    for index in range(n_clients or ctx.task.n_clients):
        module = ClientCheckpointAndStateModule(post_aggregation=[
            BestLossTorchModuleCheckpointer(str(ctx.run_dir), f"client_{index}_best_model.pkl"),
            LatestTorchModuleCheckpointer(str(ctx.run_dir), f"client_{index}_last_model.pkl")]) if ctx.spec.checkpoint else None
        client = cls(Path(ctx.spec.data_dir), ctx.task.metrics(), ctx.device, client_name=f"client_{index}",
                     checkpoint_and_state_module=module, **client_kwargs)
        client.ctx, client.client_index, client.model_factory = ctx, index, model_factory
        if customise is not None:
            customise(client)
        clients.append(client)
    return clients


def dict_optimizers(client: Any, parts: dict[str, Callable[[Any], nn.Module]]) -> None:
    client.get_optimizer = lambda config: {key: client.make_optimizer(select(client).parameters()) for key, select in parts.items()}


def _initial_parameters(ctx: MethodContext, factory: Callable[[], nn.Module]) -> tuple[nn.Module, Any]:
    torch.manual_seed(ctx.spec.seed)
    template = factory()
    return template, ndarrays_to_parameters([v.detach().clone() for v in template.state_dict().values()])


def full_exchange_server(ctx: MethodContext, strategy: Any, template: nn.Module | None = None, fn: Any = None) -> FullExchangeServer:
    return FullExchangeServer(SimpleClientManager(), {"n_server_rounds": ctx.spec.rounds}, strategy, model=template if ctx.spec.checkpoint else None,
                              checkpoint_dir=ctx.run_dir if ctx.spec.checkpoint else None, on_init_parameters_config_fn=fn or config_fn(ctx))


def personal_server(ctx: MethodContext, strategy: Any, server_cls: type = PersonalServer, fn: Any = None) -> Any:
    return make_personal(server_cls)(SimpleClientManager(), {"n_server_rounds": ctx.spec.rounds}, strategy,
                                     on_init_parameters_config_fn=fn or config_fn(ctx))


def constraint_strategy(ctx: MethodContext, adaptive: bool = False) -> Any:
    from fl4health_b200.strategies.fedavg_with_adaptive_constraint import FedAvgWithAdaptiveConstraint

    return FedAvgWithAdaptiveConstraint(initial_parameters=None, initial_loss_weight=ctx.spec.lam, adapt_loss_weight=adaptive,
                                        loss_weight_delta=ctx.spec.lam_delta, loss_weight_patience=ctx.spec.lam_patience, **strategy_kwargs(ctx))


# ------------------------------------------------------------------------------------------------- global-model methods
@method("fedavg")
def fedavg(ctx: MethodContext) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.basic_client import BasicClient

    template, _ = _initial_parameters(ctx, ctx.task.plain)
    return full_exchange_server(ctx, BasicFedAvg(**strategy_kwargs(ctx)), template), make_clients(ctx, BasicClient, ctx.task.plain)


@method("central")
def central(ctx: MethodContext) -> tuple[Any, list[Any]]:
    """All data in one place: the usual upper baseline (``research/flamby/*/central``)."""
    from fl4health_b200.clients.basic_client import BasicClient

    template, _ = _initial_parameters(ctx, ctx.task.plain)
    server = full_exchange_server(ctx, BasicFedAvg(**strategy_kwargs(ctx, n_clients=1)), template)
    return server, make_clients(ctx, BasicClient, ctx.task.plain, n_clients=1, mixin=PooledDataMixin)


@method("local")
def local(ctx: MethodContext) -> tuple[Any, list[Any]]:
    """Every client trains alone (``research/flamby/*/local``): the server's aggregate is received and ignored after
    the first round, so the protocol — and therefore the logging / checkpointing / selection code — is unchanged."""
    from fl4health_b200.clients.basic_client import BasicClient

    class LocalOnlyClient(BasicClient):
        def set_parameters(self, parameters: Any, config: Config, fitting_round: bool) -> None:
            if not getattr(self, "_initial_weights_received", False):
                super().set_parameters(parameters, config, fitting_round)
                self._initial_weights_received = True

    return personal_server(ctx, BasicFedAvg(**strategy_kwargs(ctx))), make_clients(ctx, LocalOnlyClient, ctx.task.plain)


def _fedopt(ctx: MethodContext, strategy_name: str) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.basic_client import BasicClient
    from fl4health_b200.strategies import fedopt

    template, initial = _initial_parameters(ctx, ctx.task.plain)
    strategy = getattr(fedopt, strategy_name)(initial_parameters=initial, eta=ctx.spec.server_lr, **strategy_kwargs(ctx))
    return full_exchange_server(ctx, strategy, template), make_clients(ctx, BasicClient, ctx.task.plain)


@method("fedadam")
def fedadam(ctx: MethodContext) -> tuple[Any, list[Any]]:
    return _fedopt(ctx, "FedAdam")


@method("fedyogi")
def fedyogi(ctx: MethodContext) -> tuple[Any, list[Any]]:
    return _fedopt(ctx, "FedYogi")


@method("fedprox")
def fedprox(ctx: MethodContext) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.fed_prox_client import FedProxClient
    from fl4health_b200.servers.adaptive_constraint_servers.fedprox_server import FedProxServer

    return personal_server(ctx, constraint_strategy(ctx), FedProxServer), make_clients(ctx, FedProxClient, ctx.task.plain)


@method("adaptive_fedprox")
def adaptive_fedprox(ctx: MethodContext) -> tuple[Any, list[Any]]:
    """``research/cifar10/adaptive_pfl/fedprox``: μ adapted from the aggregated training loss."""
    from fl4health_b200.clients.fed_prox_client import FedProxClient
    from fl4health_b200.servers.adaptive_constraint_servers.fedprox_server import FedProxServer

    return personal_server(ctx, constraint_strategy(ctx, adaptive=True), FedProxServer), make_clients(ctx, FedProxClient, ctx.task.plain)


@method("scaffold")
def scaffold(ctx: MethodContext) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.scaffold_client import ScaffoldClient
    from fl4health_b200.servers.scaffold_server import ScaffoldServer
    from fl4health_b200.strategies.scaffold import Scaffold

    template, initial = _initial_parameters(ctx, ctx.task.plain)
    kwargs = {k: v for k, v in strategy_kwargs(ctx).items() if k not in ("min_fit_clients", "min_evaluate_clients")}
    strategy = Scaffold(initial_parameters=initial, model=template, learning_rate=ctx.spec.server_lr_scaffold, **kwargs)
    server = make_personal(ScaffoldServer)(SimpleClientManager(), {"n_server_rounds": ctx.spec.rounds}, strategy)
    return server, make_clients(ctx, ScaffoldClient, ctx.task.plain)


@method("moon")
def moon(ctx: MethodContext) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.moon_client import MoonClient
    from fl4health_b200.model_bases.moon_base import MoonModel

    factory = lambda: MoonModel(ctx.task.features(), ctx.task.head())  # noqa: E731
    clients = make_clients(ctx, MoonClient, factory, contrastive_weight=ctx.spec.lam)
    return personal_server(ctx, BasicFedAvg(**strategy_kwargs(ctx))), clients


# ------------------------------------------------------------------------------------------------- personalised methods
@method("fedper")
def fedper(ctx: MethodContext) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.fedper_client import FedPerClient
    from fl4health_b200.model_bases.sequential_split_models import SequentiallySplitExchangeBaseModel

    factory = lambda: SequentiallySplitExchangeBaseModel(ctx.task.features(), ctx.task.head())  # noqa: E731
    return personal_server(ctx, BasicFedAvg(**strategy_kwargs(ctx))), make_clients(ctx, FedPerClient, factory)


@method("apfl")
def apfl(ctx: MethodContext) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.apfl_client import ApflClient
    from fl4health_b200.model_bases.apfl_base import ApflModule

    def customise(client: Any) -> None:
        dict_optimizers(client, {"local": lambda c: c.model.local_model, "global": lambda c: c.model.global_model})

    factory = lambda: ApflModule(ctx.task.plain(), alpha_lr=ctx.spec.alpha_lr)  # noqa: E731
    return personal_server(ctx, BasicFedAvg(**strategy_kwargs(ctx))), make_clients(ctx, ApflClient, factory, customise)


def _parallel(ctx: MethodContext, cls: type) -> Callable[[], nn.Module]:
    return lambda: cls(ctx.task.features(), ctx.task.features(), ctx.task.parallel_head())


@method("fenda")
def fenda(ctx: MethodContext) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.fenda_client import FendaClient
    from fl4health_b200.model_bases.fenda_base import FendaModel

    return personal_server(ctx, BasicFedAvg(**strategy_kwargs(ctx))), make_clients(ctx, FendaClient, _parallel(ctx, FendaModel))


@method("perfcl")
def perfcl(ctx: MethodContext) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.perfcl_client import PerFclClient
    from fl4health_b200.model_bases.perfcl_base import PerFclModel

    clients = make_clients(ctx, PerFclClient, _parallel(ctx, PerFclModel), global_feature_contrastive_loss_weight=ctx.spec.lam,
                           local_feature_contrastive_loss_weight=ctx.spec.lam)
    return personal_server(ctx, BasicFedAvg(**strategy_kwargs(ctx))), clients


def _ditto_like(ctx: MethodContext, client_cls: type, server_cls: type, twin: bool, adaptive: bool = False, **client_kwargs: Any) -> tuple[Any, list[Any]]:
    def customise(client: Any) -> None:
        if twin:
            dict_optimizers(client, {"global": lambda c: c.global_model, "local": lambda c: c.model})

    clients = make_clients(ctx, client_cls, ctx.task.plain, customise, **client_kwargs)
    return personal_server(ctx, constraint_strategy(ctx, adaptive), server_cls), clients


@method("ditto")
def ditto(ctx: MethodContext) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.ditto_client import DittoClient
    from fl4health_b200.servers.adaptive_constraint_servers.ditto_server import DittoServer

    return _ditto_like(ctx, DittoClient, DittoServer, twin=True)


@method("adaptive_ditto")
def adaptive_ditto(ctx: MethodContext) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.ditto_client import DittoClient
    from fl4health_b200.servers.adaptive_constraint_servers.ditto_server import DittoServer

    return _ditto_like(ctx, DittoClient, DittoServer, twin=True, adaptive=True)


@method("mr_mtl")
def mr_mtl(ctx: MethodContext) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.mr_mtl_client import MrMtlClient
    from fl4health_b200.servers.adaptive_constraint_servers.mrmtl_server import MrMtlServer

    return _ditto_like(ctx, MrMtlClient, MrMtlServer, twin=False)


@method("adaptive_mr_mtl")
def adaptive_mr_mtl(ctx: MethodContext) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.mr_mtl_client import MrMtlClient
    from fl4health_b200.servers.adaptive_constraint_servers.mrmtl_server import MrMtlServer

    return _ditto_like(ctx, MrMtlClient, MrMtlServer, twin=False, adaptive=True)


def _mkmmd_kwargs(ctx: MethodContext) -> dict[str, Any]:
    return dict(mkmmd_loss_weight=ctx.spec.mmd_weight, feature_extraction_layers=list(ctx.task.feature_layers),
                beta_global_update_interval=ctx.spec.beta_update_interval, num_accumulating_batches=2)


def _deep_mmd_kwargs(ctx: MethodContext) -> dict[str, Any]:
    width = ctx.task.features()(ctx.task.client_data(0, ctx.spec)[0].data[:2]).shape[1]
    return dict(deep_mmd_loss_weight=ctx.spec.mmd_weight, feature_extraction_layers_with_size={layer: int(width) for layer in ctx.task.feature_layers})


@method("ditto_mkmmd")
def ditto_mkmmd(ctx: MethodContext) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.mkmmd_clients import DittoMkMmdClient
    from fl4health_b200.servers.adaptive_constraint_servers.ditto_server import DittoServer

    return _ditto_like(ctx, DittoMkMmdClient, DittoServer, twin=True, **_mkmmd_kwargs(ctx))


@method("mr_mtl_mkmmd")
def mr_mtl_mkmmd(ctx: MethodContext) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.mkmmd_clients import MrMtlMkMmdClient
    from fl4health_b200.servers.adaptive_constraint_servers.mrmtl_server import MrMtlServer

    return _ditto_like(ctx, MrMtlMkMmdClient, MrMtlServer, twin=False, **_mkmmd_kwargs(ctx))


@method("ditto_deep_mmd")
def ditto_deep_mmd(ctx: MethodContext) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.deep_mmd_clients import DittoDeepMmdClient
    from fl4health_b200.servers.adaptive_constraint_servers.ditto_server import DittoServer

    return _ditto_like(ctx, DittoDeepMmdClient, DittoServer, twin=True, **_deep_mmd_kwargs(ctx))


@method("mr_mtl_deep_mmd")
def mr_mtl_deep_mmd(ctx: MethodContext) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.deep_mmd_clients import MrMtlDeepMmdClient
    from fl4health_b200.servers.adaptive_constraint_servers.mrmtl_server import MrMtlServer

    return _ditto_like(ctx, MrMtlDeepMmdClient, MrMtlServer, twin=False, **_deep_mmd_kwargs(ctx))


@method("fenda_ditto")
def fenda_ditto(ctx: MethodContext) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.fenda_ditto_client import FendaDittoClient
    from fl4health_b200.model_bases.fenda_base import FendaModel
    from fl4health_b200.model_bases.sequential_split_models import SequentiallySplitModel
    from fl4health_b200.servers.adaptive_constraint_servers.ditto_server import DittoServer

    def customise(client: Any) -> None:
        client.get_global_model = lambda cfg: SequentiallySplitModel(ctx.task.features(), ctx.task.head())
        dict_optimizers(client, {"global": lambda c: c.global_model, "local": lambda c: c.model})

    clients = make_clients(ctx, FendaDittoClient, _parallel(ctx, FendaModel), customise)
    return personal_server(ctx, constraint_strategy(ctx), DittoServer), clients


# ----------------------------------------------------------------------------------- generalisation-adjusted aggregation
@method("fedavg_ga")
def fedavg_ga(ctx: MethodContext) -> tuple[Any, list[Any]]:
    """``research/cifar10/fed_dgga_pfl``: FedAvg whose aggregation weights follow each client's generalisation gap."""
    from fl4health_b200.client_managers.fixed_sampling_client_manager import FixedSamplingClientManager
    from fl4health_b200.clients.basic_client import BasicClient
    from fl4health_b200.strategies.feddg_ga import FedDgGa

    fn = config_fn(ctx, evaluate_after_fit=True, pack_losses_with_val_metrics=True)
    kwargs = {k: v for k, v in strategy_kwargs(ctx, fn).items() if not k.startswith("min_")}
    server = make_personal(PersonalServer)(FixedSamplingClientManager(), {"n_server_rounds": ctx.spec.rounds}, FedDgGa(**kwargs),
                                           on_init_parameters_config_fn=fn)
    return server, make_clients(ctx, BasicClient, ctx.task.plain)


# ------------------------------------------------------------------------------------------------------ partial exchange
def _partial_exchange(ctx: MethodContext, exchanger: Callable[[], Any], strategy: Any) -> tuple[Any, list[Any]]:
    from fl4health_b200.clients.partial_weight_exchange_client import PartialWeightExchangeClient

    def customise(client: Any) -> None:
        client.store_initial_model = True
        client.get_parameter_exchanger = lambda cfg: exchanger()

    return personal_server(ctx, strategy), make_clients(ctx, PartialWeightExchangeClient, ctx.task.plain, customise)


@method("dynamic_layer")
def dynamic_layer(ctx: MethodContext) -> tuple[Any, list[Any]]:
    """``research/ag_news/dynamic_layer_exchange``: only the layers that drifted most are sent."""
    from fl4health_b200.parameter_exchange.layer_exchanger import DynamicLayerExchanger
    from fl4health_b200.parameter_exchange.parameter_selection_criteria import LayerSelectionFunctionConstructor
    from fl4health_b200.strategies.fedavg_dynamic_layer import FedAvgDynamicLayer

    select = LayerSelectionFunctionConstructor(1e-9, ctx.spec.exchange_fraction, normalize=True).select_by_percentage
    return _partial_exchange(ctx, lambda: DynamicLayerExchanger(select()), FedAvgDynamicLayer(**strategy_kwargs(ctx)))


@method("sparse_coo")
def sparse_coo(ctx: MethodContext) -> tuple[Any, list[Any]]:
    """``research/ag_news/sparse_tensor_exchange``: element-wise top-k exchange in COO form."""
    from fl4health_b200.parameter_exchange.parameter_selection_criteria import largest_final_magnitude_scores
    from fl4health_b200.parameter_exchange.sparse_coo_parameter_exchanger import SparseCooParameterExchanger
    from fl4health_b200.strategies.fedavg_sparse_coo_tensor import FedAvgSparseCooTensor

    exchanger = lambda: SparseCooParameterExchanger(ctx.spec.exchange_fraction, largest_final_magnitude_scores)  # noqa: E731
    return _partial_exchange(ctx, exchanger, FedAvgSparseCooTensor(**strategy_kwargs(ctx)))
