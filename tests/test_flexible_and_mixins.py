"""FlexibleClient API + mixins: FedProx / Ditto / MR-MTL obtained by mixing into a user FlexibleClient must behave like
the dedicated client classes (same wire protocol, same servers)."""

import warnings
from pathlib import Path

import pytest
import torch

from fl4health_b200.clients.flexible import FlexibleClient
from fl4health_b200.engine.fused_optim import _FlatOptimizer
from fl4health_b200.metrics import Accuracy
from fl4health_b200.metrics.metric_aggregation import evaluate_metrics_aggregation_fn, fit_metrics_aggregation_fn
from fl4health_b200.mixins import PersonalizedMode, apply_adaptive_drift_to_client, make_it_personal
from fl4health_b200.mixins.adaptive_drift_constrained import AdaptiveDriftConstrainedMixin
from fl4health_b200.servers.adaptive_constraint_servers.ditto_server import DittoServer
from fl4health_b200.servers.adaptive_constraint_servers.fedprox_server import FedProxServer
from fl4health_b200.servers.adaptive_constraint_servers.mrmtl_server import MrMtlServer
from fl4health_b200.servers.base_server import FlServer
from fl4health_b200.servers.client_manager import SimpleClientManager
from fl4health_b200.simulation import run_simulation
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg
from fl4health_b200.strategies.fedavg_with_adaptive_constraint import FedAvgWithAdaptiveConstraint
from fl4health_b200.utils.random import set_all_random_seeds
from tests.helpers import SyntheticCifarMixin, TinyNet, fit_config_fn


class MyFlexibleClient(SyntheticCifarMixin, FlexibleClient):
    model_fn = staticmethod(TinyNet)
    momentum = 0.0


def _clients(cls, k: int = 2):
    out = []
    for idx in range(k):
        client = cls(Path("."), [Accuracy()], torch.device("cpu"), client_name=f"f{idx}")
        client.seed = idx
        out.append(client)
    return out


def _common():
    cfg = fit_config_fn(local_steps=4)
    return dict(min_fit_clients=2, min_evaluate_clients=2, min_available_clients=2, on_fit_config_fn=cfg,
                on_evaluate_config_fn=cfg, fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
                evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn), cfg


def test_flexible_client_matches_basic_client_training() -> None:
    from tests.helpers import make_mixed_clients
    from fl4health_b200.clients.basic_client import BasicClient

    def run(clients):
        set_all_random_seeds(9)
        common, cfg = _common()
        server = FlServer(SimpleClientManager(), {"n_server_rounds": 2}, BasicFedAvg(**common), on_init_parameters_config_fn=cfg)
        return run_simulation(server, clients, 2)

    set_all_random_seeds(9)
    h_flex = run(_clients(MyFlexibleClient))
    set_all_random_seeds(9)
    h_basic = run(make_mixed_clients(BasicClient, 2, model_fn=staticmethod(TinyNet), momentum=0.0))
    for (_, a), (_, b) in zip(h_flex.losses_distributed, h_basic.losses_distributed):
        assert a == pytest.approx(b, rel=1e-5)


def test_overriding_legacy_hooks_warns() -> None:
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")

        class Legacy(FlexibleClient):
            def predict(self, input):  # noqa: ANN001, ANN201
                return super().predict(input)

    assert any("predict_with_model" in str(w.message) for w in caught)


def test_adaptive_drift_mixin_is_fedprox() -> None:
    set_all_random_seeds(11)
    cls = apply_adaptive_drift_to_client(MyFlexibleClient)
    assert cls.__name__ == "AdaptiveDriftMyFlexibleClient" and issubclass(cls, AdaptiveDriftConstrainedMixin)
    clients = _clients(cls)
    common, cfg = _common()
    strategy = FedAvgWithAdaptiveConstraint(initial_parameters=None, initial_loss_weight=0.1, adapt_loss_weight=True,
                                            loss_weight_delta=0.05, loss_weight_patience=1, **common)
    server = FedProxServer(SimpleClientManager(), {"n_server_rounds": 3}, strategy, on_init_parameters_config_fn=cfg)
    history = run_simulation(server, clients, 3)
    losses = [v for _, v in history.losses_distributed]
    assert losses[-1] < losses[0]
    opt = clients[0].optimizers["global"]
    assert isinstance(opt, _FlatOptimizer) and opt.anchor is not None  # penalty gradient applied inside the optimizer
    assert strategy.loss_weight != 0.1


def test_make_it_personal_ditto() -> None:
    set_all_random_seeds(12)
    cls = make_it_personal(MyFlexibleClient, PersonalizedMode.DITTO)
    assert cls.__name__ == "DittoMyFlexibleClient"
    clients = _clients(cls)
    common, cfg = _common()
    strategy = FedAvgWithAdaptiveConstraint(initial_parameters=None, initial_loss_weight=0.5, **common)
    server = DittoServer(SimpleClientManager(), {"n_server_rounds": 2}, strategy, on_init_parameters_config_fn=cfg)
    history = run_simulation(server, clients, 2)
    g0, g1 = clients[0].global_model.state_dict(), clients[1].global_model.state_dict()
    assert all(torch.equal(g0[k], g1[k]) for k in g0)
    p0, p1 = clients[0].model.state_dict(), clients[1].model.state_dict()
    assert any(not torch.equal(p0[k], p1[k]) for k in p0)
    assert set(clients[0].optimizers) == {"local", "global"}
    assert clients[0].optimizers["global"].param_groups[0]["lr"] == clients[0].optimizers["local"].param_groups[0]["lr"]
    assert "val - local-prediction - accuracy" in history.metrics_distributed
    assert "val - global-prediction - accuracy" in history.metrics_distributed


def test_make_it_personal_mr_mtl() -> None:
    set_all_random_seeds(13)
    clients = _clients(make_it_personal(MyFlexibleClient, PersonalizedMode.MR_MTL))
    common, cfg = _common()
    strategy = FedAvgWithAdaptiveConstraint(initial_parameters=None, initial_loss_weight=0.5, **common)
    server = MrMtlServer(SimpleClientManager(), {"n_server_rounds": 2}, strategy, on_init_parameters_config_fn=cfg)
    run_simulation(server, clients, 2)
    p0, p1 = clients[0].model.state_dict(), clients[1].model.state_dict()
    assert any(not torch.equal(p0[k], p1[k]) for k in p0)
    a0, a1 = clients[0].initial_global_model.state_dict(), clients[1].initial_global_model.state_dict()
    assert all(torch.equal(a0[k], a1[k]) for k in a0)
    with pytest.raises(ValueError):
        make_it_personal(MyFlexibleClient, "nope")  # type: ignore[arg-type]
