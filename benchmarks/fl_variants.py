"""Federation variants for ``bench.py --config``: the BASELINE.json configurations beyond plain FedAvg, on the same
synthetic CIFAR-10-shaped data and the same timing harness.

    scaffold_fedprox  -> "scaffold" (ScaffoldClient / Scaffold / ScaffoldServer, vanilla SGD) and
                         "fedprox"  (FedProxClient / FedAvgWithAdaptiveConstraint, mu = 0.1 fixed)
    fedper_ditto_dp   -> "fedper"   (ResNet-18 features exchanged, head personal), "ditto" (global + personal twin, lambda = 0.1)
                         and "dp_sgd" (InstanceLevelDpClient on the reference's CIFAR LeNet, sigma = 1.0, C = 5.0)

``build(variant, ...)`` returns ``(client, server)`` ready for ``build_spmd_federation``; the client class is the
algorithm's client mixed with ``hooks`` (the harness' data / criterion / optimizer factories).
"""

from __future__ import annotations

from pathlib import Path
from typing import Any

import torch
from torch import nn

from fl4health_b200.common.typing import ndarrays_to_parameters
from fl4health_b200.metrics import Accuracy
from fl4health_b200.metrics.metric_aggregation import evaluate_metrics_aggregation_fn, fit_metrics_aggregation_fn
from fl4health_b200.models import resnet18_cifar
from fl4health_b200.servers.client_manager import SimpleClientManager

GROUPS = {"cifar_fedavg": ["fedavg"], "scaffold_fedprox": ["scaffold", "fedprox"], "fedper_ditto_dp": ["fedper", "ditto", "dp_sgd"]}


def _rng_devices() -> list[int]:
    """Devices whose generators ``torch.manual_seed`` would rewind: forked together with the CPU generator so that a
    seeded model construction leaves the client's random streams (mask sampling, dropout, DP noise) where they were."""
    return [torch.cuda.current_device()] if torch.cuda.is_available() and torch.cuda.is_initialized() else []


class _ResNetFeatures(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.net = resnet18_cifar()
        self.net.fc = nn.Identity()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.net.forward_features(x)


class LeNetCifar(nn.Module):
    """The reference's ``Net`` (examples/models/cnn_model.py:6-39): conv5(3->6)-pool-conv5(6->16)-pool-fc120-fc84-fc10."""

    def __init__(self) -> None:
        super().__init__()
        self.conv1, self.conv2 = nn.Conv2d(3, 6, 5), nn.Conv2d(6, 16, 5)
        self.pool = nn.MaxPool2d(2, 2)
        self.fc1, self.fc2, self.fc3 = nn.Linear(16 * 5 * 5, 120), nn.Linear(120, 84), nn.Linear(84, 10)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.pool(torch.relu(self.conv1(x)))
        x = self.pool(torch.relu(self.conv2(x)))
        x = torch.flatten(x, 1)
        return self.fc3(torch.relu(self.fc2(torch.relu(self.fc1(x)))))


def _strategy_options(world: int, config_fn: Any) -> dict[str, Any]:
    return dict(min_fit_clients=world, min_evaluate_clients=world, min_available_clients=world, on_fit_config_fn=config_fn,
                on_evaluate_config_fn=config_fn, fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
                evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn)


def _initial_parameters(model: nn.Module) -> Any:
    return ndarrays_to_parameters([v.detach().clone() for v in model.state_dict().values()])


def build(variant: str, hooks: type, ctx: Any, engine: Any, rounds: int, local_steps: int, batch_size: int) -> tuple[Any, Any]:
    from fl4health_b200.servers.base_server import FlServer
    from fl4health_b200.strategies.basic_fedavg import BasicFedAvg

    world, device = ctx.world_size, ctx.device
    lr = 0.01

    def config_fn(server_round: int) -> dict:
        return {"current_server_round": server_round, "local_steps": local_steps, "batch_size": batch_size}

    fl_config = {"n_server_rounds": rounds, "local_steps": local_steps}
    options = _strategy_options(world, config_fn)

    def make_client(algorithm: type, model_factory: Any, optimizer_factory: Any = None, **extra: Any) -> Any:
        namespace: dict[str, Any] = {"get_model": lambda self, config: model_factory()}
        if optimizer_factory is not None:
            namespace["get_optimizer"] = optimizer_factory
        cls = type(f"Bench{algorithm.__name__}", (hooks, algorithm), namespace)
        return cls(Path("."), [Accuracy()], device, client_name=f"rank{ctx.rank}", engine_options=engine, **extra)

    def seeded(factory: Any) -> Any:
        def make() -> nn.Module:
            with torch.random.fork_rng(devices=_rng_devices()):  # same initialisation everywhere, ambient random stream untouched
                torch.manual_seed(1234)
                return factory()
        return make

    if variant == "fedavg":
        from fl4health_b200.clients.basic_client import BasicClient

        client = make_client(BasicClient, seeded(resnet18_cifar))
        return client, FlServer(SimpleClientManager(), fl_config, BasicFedAvg(**options), on_init_parameters_config_fn=config_fn,
                                accept_failures=False)
    if variant == "fedprox":
        from fl4health_b200.clients.fed_prox_client import FedProxClient
        from fl4health_b200.servers.adaptive_constraint_servers.fedprox_server import FedProxServer
        from fl4health_b200.strategies.fedavg_with_adaptive_constraint import FedAvgWithAdaptiveConstraint

        strategy = FedAvgWithAdaptiveConstraint(initial_parameters=None, initial_loss_weight=0.1, adapt_loss_weight=False, **options)
        client = make_client(FedProxClient, seeded(resnet18_cifar))
        return client, FedProxServer(SimpleClientManager(), fl_config, strategy, on_init_parameters_config_fn=config_fn, accept_failures=False)
    if variant == "scaffold":
        from fl4health_b200.clients.scaffold_client import ScaffoldClient
        from fl4health_b200.servers.scaffold_server import ScaffoldServer
        from fl4health_b200.strategies.scaffold import Scaffold

        template = seeded(resnet18_cifar)()
        keep = {k: v for k, v in options.items() if k not in ("min_fit_clients", "min_evaluate_clients")}
        strategy = Scaffold(initial_parameters=_initial_parameters(template), model=template, learning_rate=1.0, **keep)
        strategy.min_fit_clients = strategy.min_evaluate_clients = world  # the constructor fixes them at 2 (as the reference's does)

        def vanilla_sgd(self: Any, config: dict) -> Any:  # SCAFFOLD requires plain SGD (scaffold_client.py:293)
            return torch.optim.SGD(self.model.parameters(), lr=lr)

        client = make_client(ScaffoldClient, seeded(resnet18_cifar), vanilla_sgd)
        client.learning_rate = lr
        return client, ScaffoldServer(SimpleClientManager(), fl_config, strategy, on_init_parameters_config_fn=config_fn, accept_failures=False)
    if variant == "fedper":
        from fl4health_b200.clients.fedper_client import FedPerClient
        from fl4health_b200.model_bases.sequential_split_models import SequentiallySplitExchangeBaseModel

        factory = seeded(lambda: SequentiallySplitExchangeBaseModel(_ResNetFeatures(), nn.Linear(512, 10)))
        client = make_client(FedPerClient, factory)
        return client, FlServer(SimpleClientManager(), fl_config, BasicFedAvg(**options), on_init_parameters_config_fn=config_fn,
                                accept_failures=False)
    if variant == "ditto":
        from fl4health_b200.clients.ditto_client import DittoClient
        from fl4health_b200.servers.adaptive_constraint_servers.ditto_server import DittoServer
        from fl4health_b200.strategies.fedavg_with_adaptive_constraint import FedAvgWithAdaptiveConstraint

        def twin_optimizers(self: Any, config: dict) -> dict:
            return {"global": torch.optim.SGD(self.global_model.parameters(), lr=lr, momentum=0.9),
                    "local": torch.optim.SGD(self.model.parameters(), lr=lr, momentum=0.9)}

        strategy = FedAvgWithAdaptiveConstraint(initial_parameters=None, initial_loss_weight=0.1, adapt_loss_weight=False, **options)
        client = make_client(DittoClient, seeded(resnet18_cifar), twin_optimizers)
        return client, DittoServer(SimpleClientManager(), fl_config, strategy, on_init_parameters_config_fn=config_fn, accept_failures=False)
    if variant == "dp_sgd":
        from fl4health_b200.client_managers.poisson_sampling_manager import PoissonSamplingClientManager
        from fl4health_b200.clients.instance_level_dp_client import InstanceLevelDpClient
        from fl4health_b200.servers.instance_level_dp_server import InstanceLevelDpServer

        def dp_config(server_round: int) -> dict:
            return {**config_fn(server_round), "clipping_bound": 5.0, "noise_multiplier": 1.0}

        dp_options = {**_strategy_options(world, dp_config), "fraction_fit": 1.0, "fraction_evaluate": 1.0}
        client = make_client(InstanceLevelDpClient, seeded(LeNetCifar))
        server = InstanceLevelDpServer(PoissonSamplingClientManager(), fl_config, noise_multiplier=1.0, batch_size=batch_size,
                                       num_server_rounds=rounds, strategy=BasicFedAvg(**dp_options), local_steps=local_steps,
                                       on_init_parameters_config_fn=dp_config, accept_failures=False)
        return client, server
    raise ValueError(f"unknown variant {variant}")
