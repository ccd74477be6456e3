"""FENDA / constrained FENDA / PerFCL / FENDA+Ditto / GPFL clients: short CPU federations + invariants."""

import torch
from torch import nn

from fl4health_b200.clients.constrained_fenda_client import ConstrainedFendaClient
from fl4health_b200.clients.fenda_client import FendaClient
from fl4health_b200.clients.fenda_ditto_client import FendaDittoClient
from fl4health_b200.clients.gpfl_client import GpflClient
from fl4health_b200.clients.perfcl_client import PerFclClient
from fl4health_b200.losses.fenda_loss_config import (
    ConstrainedFendaLossContainer,
    CosineSimilarityLossContainer,
    MoonContrastiveLossContainer,
    PerFclLossContainer,
)
from fl4health_b200.metrics.metric_aggregation import evaluate_metrics_aggregation_fn, fit_metrics_aggregation_fn
from fl4health_b200.model_bases.fenda_base import FendaModel, FendaModelWithFeatureState
from fl4health_b200.model_bases.gpfl_base import GpflModel
from fl4health_b200.model_bases.parallel_split_models import ParallelFeatureJoinMode, ParallelSplitHeadModule
from fl4health_b200.model_bases.perfcl_base import PerFclModel
from fl4health_b200.model_bases.sequential_split_models import SequentiallySplitModel
from fl4health_b200.servers.adaptive_constraint_servers.ditto_server import DittoServer
from fl4health_b200.servers.base_server import FlServer
from fl4health_b200.servers.client_manager import SimpleClientManager
from fl4health_b200.simulation import run_simulation
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg
from fl4health_b200.strategies.fedavg_with_adaptive_constraint import FedAvgWithAdaptiveConstraint
from fl4health_b200.utils.random import set_all_random_seeds
from tests.helpers import fit_config_fn, make_mixed_clients

CPU = torch.device("cpu")


class Extractor(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.conv = nn.Conv2d(3, 4, 3, padding=1)

    def forward(self, x):
        return torch.flatten(torch.nn.functional.adaptive_avg_pool2d(torch.relu(self.conv(x)), 2), 1)  # [B,16]


class JoinHead(ParallelSplitHeadModule):
    def __init__(self) -> None:
        super().__init__(ParallelFeatureJoinMode.CONCATENATE)
        self.fc = nn.Linear(32, 10)

    def parallel_output_join(self, local_tensor, global_tensor):
        return torch.cat([local_tensor, global_tensor], dim=1)

    def head_forward(self, input_tensor):
        return self.fc(input_tensor)


class PlainHead(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.fc = nn.Linear(16, 10)

    def forward(self, x):
        return self.fc(x)


def _common(cfg=None):
    cfg = cfg or fit_config_fn()
    return dict(min_fit_clients=2, min_evaluate_clients=2, min_available_clients=2, on_fit_config_fn=cfg,
                on_evaluate_config_fn=cfg, fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
                evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn)


def _run(clients, rounds=3, strategy=None, server_cls=FlServer):
    server = server_cls(SimpleClientManager(), {"n_server_rounds": rounds}, strategy or BasicFedAvg(**_common()),
                        on_init_parameters_config_fn=fit_config_fn())
    return run_simulation(server, clients, rounds)


def _shared(clients, prefix, attr="model"):
    s0, s1 = getattr(clients[0], attr).state_dict(), getattr(clients[1], attr).state_dict()
    same = all(torch.equal(s0[k], s1[k]) for k in s0 if k.startswith(prefix))
    other_differs = any(not torch.equal(s0[k], s1[k]) for k in s0 if not k.startswith(prefix))
    return same, other_differs


def test_fenda_exchanges_only_global_extractor() -> None:
    set_all_random_seeds(41)
    clients = make_mixed_clients(FendaClient, 2, model_fn=staticmethod(lambda: FendaModel(Extractor(), Extractor(), JoinHead())))
    history = _run(clients, 2)
    same, differs = _shared(clients, "second_feature_extractor.")
    assert same and differs and len(history.losses_distributed) == 2


def test_constrained_fenda_all_losses_active() -> None:
    set_all_random_seeds(42)
    container = lambda: ConstrainedFendaLossContainer(  # noqa: E731
        PerFclLossContainer(CPU, 0.5, 0.5), CosineSimilarityLossContainer(CPU, 0.1), MoonContrastiveLossContainer(CPU, 0.5))
    clients = make_mixed_clients(ConstrainedFendaClient, 2, model_fn=staticmethod(
        lambda: FendaModelWithFeatureState(Extractor(), Extractor(), JoinHead(), flatten_features=True)))
    for c in clients:
        c.loss_container = container()
    _run(clients, 3)
    c0 = clients[0]
    c0.model.train()
    c0.update_before_train(4)
    losses, _ = c0.train_step(torch.randn(4, 3, 32, 32), torch.randint(0, 10, (4,)))
    keys = set(losses.additional_losses)
    assert {"loss", "cos_sim_loss", "contrastive_loss", "global_feature_contrastive_loss",
            "local_feature_contrastive_loss", "total_loss"} <= keys


def test_perfcl_client() -> None:
    set_all_random_seeds(43)
    clients = make_mixed_clients(PerFclClient, 2, model_fn=staticmethod(lambda: PerFclModel(Extractor(), Extractor(), JoinHead())))
    history = _run(clients, 3)
    same, differs = _shared(clients, "second_feature_extractor.")
    assert same and differs
    assert clients[0]._all_contrastive_loss_modules_defined() and len(history.losses_distributed) == 3


def test_fenda_ditto_client() -> None:
    set_all_random_seeds(44)
    clients = make_mixed_clients(FendaDittoClient, 2)
    for c in clients:
        c.get_model = lambda config: FendaModel(Extractor(), Extractor(), JoinHead())
        c.get_global_model = lambda config: SequentiallySplitModel(Extractor(), PlainHead())
        c.get_optimizer = (lambda self, config: {"global": torch.optim.SGD(self.global_model.parameters(), lr=0.05),
                                                 "local": torch.optim.SGD(self.model.parameters(), lr=0.05)}).__get__(c)
    strategy = FedAvgWithAdaptiveConstraint(initial_parameters=None, initial_loss_weight=0.1, **_common())
    _run(clients, 2, strategy=strategy, server_cls=DittoServer)
    g0, g1 = clients[0].global_model.state_dict(), clients[1].global_model.state_dict()
    assert all(torch.equal(g0[k], g1[k]) for k in g0)
    # FENDA global extractor == Ditto global model's extractor after set_parameters
    for a, b in zip(clients[0].model.second_feature_extractor.parameters(), clients[0].global_model.base_module.parameters()):
        assert a.shape == b.shape


def test_gpfl_client() -> None:
    set_all_random_seeds(45)
    clients = make_mixed_clients(GpflClient, 2, model_fn=staticmethod(
        lambda: GpflModel(Extractor(), PlainHead(), feature_dim=16, num_classes=10)))
    for c in clients:
        c.get_optimizer = (lambda self, config: {
            "model": torch.optim.SGD(self.model.gpfl_main_module.parameters(), lr=0.05),
            "gce": torch.optim.SGD(self.model.gce.parameters(), lr=0.05),
            "cov": torch.optim.SGD(self.model.cov.parameters(), lr=0.05)}).__get__(c)
    history = _run(clients, 2)
    assert len(history.losses_distributed) == 2
    s0, s1 = clients[0].model.state_dict(), clients[1].model.state_dict()
    assert all(torch.equal(s0[k], s1[k]) for k in s0 if k.startswith(("cov.", "gce.", "gpfl_main_module.base_module.")))
    assert any(not torch.equal(s0[k], s1[k]) for k in s0 if k.startswith("gpfl_main_module.head_module."))
    assert all(g["weight_decay"] == clients[0].mu for g in clients[0].optimizers["gce"].param_groups)
    assert abs(float(clients[0].class_sample_proportion.sum()) - 1.0) < 1e-6
