"""Personalised-FL clients: Ditto, MR-MTL, APFL, FedPer, FedRep, FedBN — invariants + short e2e runs (CPU)."""

import copy

import pytest
import torch
from torch import nn

from fl4health_b200.clients.apfl_client import ApflClient
from fl4health_b200.clients.ditto_client import DittoClient
from fl4health_b200.clients.fedbn_client import FedBnClient
from fl4health_b200.clients.fedper_client import FedPerClient
from fl4health_b200.clients.fedrep_client import FedRepClient
from fl4health_b200.clients.mr_mtl_client import MrMtlClient
from fl4health_b200.metrics.metric_aggregation import evaluate_metrics_aggregation_fn, fit_metrics_aggregation_fn
from fl4health_b200.model_bases.apfl_base import ApflModule
from fl4health_b200.model_bases.fedrep_base import FedRepModel
from fl4health_b200.model_bases.sequential_split_models import SequentiallySplitExchangeBaseModel
from fl4health_b200.parameter_exchange.layer_exchanger import LayerExchangerWithExclusions
from fl4health_b200.servers.adaptive_constraint_servers.ditto_server import DittoServer
from fl4health_b200.servers.adaptive_constraint_servers.mrmtl_server import MrMtlServer
from fl4health_b200.servers.base_server import FlServer
from fl4health_b200.servers.client_manager import SimpleClientManager
from fl4health_b200.simulation import run_simulation
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg
from fl4health_b200.strategies.fedavg_with_adaptive_constraint import FedAvgWithAdaptiveConstraint
from fl4health_b200.utils.random import set_all_random_seeds
from tests.helpers import TinyNet, fit_config_fn, make_mixed_clients


def _common(cfg=None):
    cfg = cfg or fit_config_fn()
    return dict(min_fit_clients=2, min_evaluate_clients=2, min_available_clients=2, on_fit_config_fn=cfg,
                on_evaluate_config_fn=cfg, fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
                evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn)


class Base(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.conv = nn.Conv2d(3, 4, 3, padding=1)
        self.bn = nn.BatchNorm2d(4)

    def forward(self, x):
        return torch.flatten(torch.nn.functional.adaptive_avg_pool2d(torch.relu(self.bn(self.conv(x))), 4), 1)


class Head(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.fc = nn.Linear(64, 10)

    def forward(self, x):
        return self.fc(x)


def _ditto_like(client_cls, server_cls):
    set_all_random_seeds(21)

    def optimizers(self, config):
        if client_cls is DittoClient:
            return {"global": torch.optim.SGD(self.global_model.parameters(), lr=0.05),
                    "local": torch.optim.SGD(self.model.parameters(), lr=0.05)}
        return torch.optim.SGD(self.model.parameters(), lr=0.05)

    clients = make_mixed_clients(client_cls, 2, model_fn=staticmethod(TinyNet))
    for c in clients:
        c.get_optimizer = optimizers.__get__(c)
    strategy = FedAvgWithAdaptiveConstraint(initial_parameters=None, initial_loss_weight=0.5, **_common())
    server = server_cls(SimpleClientManager(), {"n_server_rounds": 2}, strategy,
                        on_init_parameters_config_fn=fit_config_fn())
    history = run_simulation(server, clients, 2)
    return clients, history


def test_ditto_keeps_personal_models_distinct_and_shares_global() -> None:
    clients, history = _ditto_like(DittoClient, DittoServer)
    g0, g1 = clients[0].global_model.state_dict(), clients[1].global_model.state_dict()
    assert all(torch.equal(g0[k], g1[k]) for k in g0)  # same aggregate after the evaluate round
    p0, p1 = clients[0].model.state_dict(), clients[1].model.state_dict()
    assert any(not torch.equal(p0[k], p1[k]) for k in p0)  # personal models differ
    assert len(history.losses_distributed) == 2
    assert "val - local - accuracy" in history.metrics_distributed and "val - global - accuracy" in history.metrics_distributed
    from fl4health_b200.engine.fused_optim import FlatSGD

    assert isinstance(clients[0].optimizers["local"], FlatSGD) and clients[0].optimizers["local"].anchor is not None
    assert clients[0].optimizers["global"].anchor is None


def test_mr_mtl_never_overwrites_personal_model() -> None:
    clients, _ = _ditto_like(MrMtlClient, MrMtlServer)
    p0, p1 = clients[0].model.state_dict(), clients[1].model.state_dict()
    assert any(not torch.equal(p0[k], p1[k]) for k in p0)
    a0, a1 = clients[0].initial_global_model.state_dict(), clients[1].initial_global_model.state_dict()
    assert all(torch.equal(a0[k], a1[k]) for k in a0)


def test_apfl_exchanges_only_global_and_adapts_alpha() -> None:
    set_all_random_seeds(22)
    clients = make_mixed_clients(ApflClient, 2, model_fn=staticmethod(lambda: ApflModule(TinyNet(), alpha_lr=0.5)))
    for c in clients:
        c.get_optimizer = (lambda self, config: {
            "local": torch.optim.SGD(self.model.local_model.parameters(), lr=0.05),
            "global": torch.optim.SGD(self.model.global_model.parameters(), lr=0.05)}).__get__(c)
    server = FlServer(SimpleClientManager(), {"n_server_rounds": 2}, BasicFedAvg(**_common()),
                      on_init_parameters_config_fn=fit_config_fn())
    history = run_simulation(server, clients, 2)
    m0, m1 = clients[0].model, clients[1].model
    assert all(torch.equal(a, b) for a, b in zip(m0.global_model.state_dict().values(), m1.global_model.state_dict().values()))
    assert any(not torch.equal(a, b) for a, b in zip(m0.local_model.parameters(), m1.local_model.parameters()))
    assert m0.alpha != 0.5 and 0.0 <= m0.alpha <= 1.0
    assert {"val - personal - accuracy", "val - global - accuracy", "val - local - accuracy"} <= set(history.metrics_distributed)


def test_apfl_alpha_gradient_flat_equals_per_layer() -> None:
    torch.manual_seed(0)
    from fl4health_b200.parallel.arena import attach_arena

    module = ApflModule(TinyNet())
    plain = copy.deepcopy(module)
    attach_arena(module)
    x, y = torch.randn(8, 3, 32, 32), torch.randint(0, 10, (8,))
    for m in (module, plain):
        with torch.no_grad():
            for p in m.local_model.parameters():
                p.add_(0.01)
        out = m(x)
        (nn.functional.cross_entropy(out["personal"], y) + nn.functional.cross_entropy(out["global"], y)).backward()
    assert module._alpha_gradient() == pytest.approx(plain._alpha_gradient(), rel=1e-4, abs=1e-6)


def test_fedper_shares_base_keeps_head() -> None:
    set_all_random_seeds(23)
    clients = make_mixed_clients(FedPerClient, 2, model_fn=staticmethod(lambda: SequentiallySplitExchangeBaseModel(Base(), Head())))
    server = FlServer(SimpleClientManager(), {"n_server_rounds": 2}, BasicFedAvg(**_common()),
                      on_init_parameters_config_fn=fit_config_fn())
    run_simulation(server, clients, 2)
    s0, s1 = clients[0].model.state_dict(), clients[1].model.state_dict()
    assert all(torch.equal(s0[k], s1[k]) for k in s0 if k.startswith("base_module."))
    assert any(not torch.equal(s0[k], s1[k]) for k in s0 if k.startswith("head_module."))


def test_fedrep_two_phase_training() -> None:
    set_all_random_seeds(24)

    def cfg(r):
        return {"current_server_round": r, "local_head_steps": 3, "local_rep_steps": 3, "batch_size": 32}

    clients = make_mixed_clients(FedRepClient, 2, model_fn=staticmethod(lambda: FedRepModel(Base(), Head())))
    for c in clients:
        c.get_optimizer = (lambda self, config: {
            "representation": torch.optim.SGD(self.model.base_module.parameters(), lr=0.05),
            "head": torch.optim.SGD(self.model.head_module.parameters(), lr=0.05)}).__get__(c)
    server = FlServer(SimpleClientManager(), {"n_server_rounds": 2}, BasicFedAvg(**_common(cfg)),
                      on_init_parameters_config_fn=cfg)
    history = run_simulation(server, clients, 2)
    fit_metrics = history.metrics_distributed_fit
    assert "head_train - prediction - accuracy" in fit_metrics and "rep_train - prediction - accuracy" in fit_metrics
    s0, s1 = clients[0].model.state_dict(), clients[1].model.state_dict()
    assert all(torch.equal(s0[k], s1[k]) for k in s0 if k.startswith("base_module."))
    assert any(not torch.equal(s0[k], s1[k]) for k in s0 if k.startswith("head_module."))


def test_fedbn_excludes_batchnorm_state() -> None:
    set_all_random_seeds(25)
    clients = make_mixed_clients(FedBnClient, 2, model_fn=staticmethod(TinyNet))
    for c in clients:
        c.get_parameter_exchanger = (lambda self, config: LayerExchangerWithExclusions(self.model, {nn.BatchNorm2d})).__get__(c)
    server = FlServer(SimpleClientManager(), {"n_server_rounds": 2}, BasicFedAvg(**_common()),
                      on_init_parameters_config_fn=fit_config_fn())
    run_simulation(server, clients, 2)
    s0, s1 = clients[0].model.state_dict(), clients[1].model.state_dict()
    assert all(torch.equal(s0[k], s1[k]) for k in s0 if not k.startswith("bn."))
    assert any(not torch.equal(s0[k], s1[k]) for k in s0 if k.startswith("bn.") and s0[k].is_floating_point())
