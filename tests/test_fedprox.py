"""FedProx / adaptive drift constraint: closed-form penalty oracles (reference: tests/clients/test_fedprox_client.py,
tests/strategies/test_fedavg_with_adaptive_constraint.py) + fused-vs-autograd equivalence + e2e."""

import copy

import numpy as np
import pytest
import torch

from fl4health_b200.clients.fed_prox_client import FedProxClient
from fl4health_b200.common.typing import ndarrays_to_parameters
from fl4health_b200.engine.fused_optim import _FlatOptimizer
from fl4health_b200.metrics.metric_aggregation import evaluate_metrics_aggregation_fn, fit_metrics_aggregation_fn
from fl4health_b200.servers.adaptive_constraint_servers.fedprox_server import FedProxServer
from fl4health_b200.servers.client_manager import SimpleClientManager
from fl4health_b200.simulation import run_simulation
from fl4health_b200.strategies.fedavg_with_adaptive_constraint import FedAvgWithAdaptiveConstraint
from fl4health_b200.utils.random import set_all_random_seeds
from tests.client_fixtures import LinearTransform, SmallMlp, build_client
from tests.helpers import fit_config_fn, make_mixed_clients


def _prime(client, mu: float) -> None:
    params = client.parameter_exchanger.push_parameters(client.model)
    packed = client.parameter_exchanger.pack_parameters(params, mu)
    client.set_parameters(packed, {"current_server_round": 1}, fitting_round=True)
    client.update_before_train(1)


def _perturb(model, delta: float) -> None:
    with torch.no_grad():
        for p in model.parameters():
            p.add_(delta)


@pytest.mark.parametrize("arena", [True, False])
def test_proximal_loss_value_and_gradient(arena: bool) -> None:
    torch.manual_seed(42)
    client = build_client(FedProxClient, SmallMlp(), arena=arena)
    _prime(client, mu=2.0)
    assert client.drift_penalty_weight == 2.0
    n_params = sum(p.numel() for p in client.model.parameters())
    assert len(client.drift_penalty_tensors) == 4
    zero = client.penalty_loss_function(client.model, client.drift_penalty_tensors, 2.0)
    assert zero.item() == 0.0
    _perturb(client.model, 0.1)
    loss = client.penalty_loss_function(client.model, client.drift_penalty_tensors, 2.0)
    assert loss.item() == pytest.approx(2.0 / 2.0 * 0.01 * n_params, rel=1e-4)
    client.optimizers["global"].zero_grad()
    loss.backward()
    for p in client.model.parameters():  # d/dw mu/2 (w - w_t)^2 = mu * 0.1
        assert torch.allclose(p.grad, torch.full_like(p, 0.2), atol=1e-5)


def test_proximal_loss_derivative_linear() -> None:
    client = build_client(FedProxClient, LinearTransform(), arena=False, d_in=2, classes=3)
    _prime(client, mu=0.1)
    _perturb(client.model, 0.1)
    loss = client.penalty_loss_function(client.model, client.drift_penalty_tensors, client.drift_penalty_weight)
    loss.backward()
    torch.testing.assert_close(client.model.linear.weight.grad, torch.full((3, 2), 0.01), atol=1e-4, rtol=1e-3)


def test_fused_penalty_step_equals_autograd_step() -> None:
    """One train step with the penalty folded into the flat optimizer == the same step through autograd."""
    torch.manual_seed(0)
    base = SmallMlp()
    fused = build_client(FedProxClient, copy.deepcopy(base), arena=True)
    plain = build_client(FedProxClient, copy.deepcopy(base), arena=False)
    assert isinstance(fused.optimizers["global"], _FlatOptimizer)
    x, y = torch.randn(16, 8), torch.randint(0, 3, (16,))
    for client in (fused, plain):
        _prime(client, mu=0.5)
        _perturb(client.model, 0.05)  # move away from the anchor so the penalty is active
        losses, _ = client.train_step(x, y)
        assert losses.additional_losses["penalty_loss"].item() == pytest.approx(
            0.5 / 2 * 0.05**2 * sum(p.numel() for p in client.model.parameters()), rel=1e-3)
    for a, b in zip(fused.model.parameters(), plain.model.parameters()):
        assert torch.allclose(a, b, atol=1e-6)


def test_adaptive_mu_rule() -> None:
    init = ndarrays_to_parameters([np.ones((2, 2), dtype=np.float32)])
    strategy = FedAvgWithAdaptiveConstraint(initial_parameters=init, initial_loss_weight=0.1, adapt_loss_weight=True,
                                            loss_weight_delta=0.05, loss_weight_patience=3)
    assert float(init.tensors[-1]) == pytest.approx(0.1)  # mu packed behind the weights
    for loss in (1.0, 0.9, 0.8):  # three non-increasing rounds -> decrease once
        strategy._maybe_update_constraint_weight_param(loss)
    assert strategy.loss_weight == pytest.approx(0.05)
    strategy._maybe_update_constraint_weight_param(0.85)  # increase -> bump up, counter reset
    assert strategy.loss_weight == pytest.approx(0.10)
    for loss in (0.8, 0.7, 0.6, 0.5, 0.4, 0.3):
        strategy._maybe_update_constraint_weight_param(loss)
    assert strategy.loss_weight == pytest.approx(0.0)  # floored at zero


def _run_fedprox(client_cls=FedProxClient, engine=None, rounds: int = 3):  # noqa: ANN001, ANN202
    set_all_random_seeds(5)
    clients = make_mixed_clients(client_cls, 2, engine=engine)
    strategy = FedAvgWithAdaptiveConstraint(
        min_fit_clients=2, min_evaluate_clients=2, min_available_clients=2, on_fit_config_fn=fit_config_fn(),
        on_evaluate_config_fn=fit_config_fn(), fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
        evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn, initial_parameters=None,
        initial_loss_weight=0.1, adapt_loss_weight=True, loss_weight_delta=0.05, loss_weight_patience=1,
    )
    server = FedProxServer(SimpleClientManager(), {"n_server_rounds": rounds}, strategy,
                           on_init_parameters_config_fn=fit_config_fn())
    history = run_simulation(server, clients, rounds)
    history.strategy = strategy
    return history, clients


def test_fedprox_end_to_end() -> None:
    history, clients = _run_fedprox()
    strategy = history.strategy
    losses = [l for _, l in history.losses_distributed]
    assert losses[-1] < losses[0]
    assert isinstance(clients[0].optimizers["global"], _FlatOptimizer)
    assert clients[0].drift_penalty_weight is not None and clients[0].drift_penalty_weight <= 0.1 + 1e-9
    assert strategy.loss_weight != 0.1  # mu was adapted
