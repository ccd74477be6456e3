"""SCAFFOLD: closed-form server/client oracles (values from the reference's tests/strategies/test_scaffold.py and
tests/clients/test_scaffold_client.py), arena fast path == per-tensor path, e2e run with BN + frozen layer."""

import copy

import numpy as np
import pytest
import torch

from fl4health_b200.clients.scaffold_client import ScaffoldClient
from fl4health_b200.common.typing import NDArrays, ndarrays_to_parameters, parameters_to_ndarrays
from fl4health_b200.engine.fused_optim import FlatSGD
from fl4health_b200.metrics.metric_aggregation import evaluate_metrics_aggregation_fn, fit_metrics_aggregation_fn
from fl4health_b200.servers.client_manager import SimpleClientManager
from fl4health_b200.servers.scaffold_server import ScaffoldServer
from fl4health_b200.simulation import run_simulation
from fl4health_b200.strategies.scaffold import Scaffold
from fl4health_b200.utils.random import set_all_random_seeds
from tests.client_fixtures import SmallMlp, build_client
from tests.helpers import TinyNet, fit_config_fn, make_mixed_clients


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def _strategy(layers, **kw):
    params = ndarrays_to_parameters([l.copy() for l in layers])
    variates = ndarrays_to_parameters([np.zeros_like(l) for l in layers])
    return Scaffold(initial_parameters=params, initial_control_variates=variates, **kw)


def test_server_aggregate_is_uniform_mean() -> None:
    layers = [np.ones(10) for _ in range(5)]
    strategy = _strategy(layers)
    clients = [NDArrays([l * (c + 1) for l in layers] + [l * (c + 1) for l in layers]) for c in range(3)]
    out = strategy.aggregate(clients)
    assert all((_np(o) == 2.0).all() for o in out)


def test_server_update_rules() -> None:
    strategy = _strategy([np.ones(10) for _ in range(5)])
    updated = strategy.compute_updated_parameters(0.1, NDArrays([np.ones(10) * 3] * 5), NDArrays([np.ones(10) * 10] * 5))
    assert all(np.allclose(_np(u), 4.0) for u in updated)  # 3 + 0.1 * 10
    strategy = _strategy([np.ones((100, 10)) * 6.0 for _ in range(10)], learning_rate=0.25)
    updated = strategy.compute_updated_weights(NDArrays([np.ones((100, 10)) * 46.0] * 10))
    assert all(np.allclose(_np(u), 16.0) for u in updated)  # 6 + 0.25 * (46 - 6)
    strategy = _strategy([np.ones((20, 10)) for _ in range(5)], fraction_fit=0.5)
    updated = strategy.compute_updated_control_variates(NDArrays([np.ones((20, 10)) * 20.0] * 5))
    assert all(np.allclose(_np(u), 10.0) for u in updated)  # 0 + 0.5 * 20


def test_client_control_variate_formula() -> None:
    client = build_client(ScaffoldClient, SmallMlp(), arena=False, lr=0.01)
    client.learning_rate = 0.01
    delta_w = NDArrays([torch.ones(4) * 2.0])
    delta_c = NDArrays([torch.ones(4) * 3.0])
    out = client.compute_updated_control_variates(5, delta_w, delta_c)
    assert torch.allclose(out[0], torch.full((4,), 3.0 + 2.0 / (5 * 0.01)))
    assert torch.allclose(client.compute_parameters_delta(delta_c, delta_w)[0], torch.ones(4))


def _packed_server_payload(client, variate_value: float):
    weights = client.parameter_exchanger.push_parameters(client.model)
    variates = NDArrays([torch.full_like(p, variate_value) for p in client.model.parameters() if p.requires_grad])
    return client.parameter_exchanger.pack_parameters(NDArrays([w.clone() for w in weights]), variates)


def test_arena_fast_path_matches_per_tensor_path() -> None:
    torch.manual_seed(3)
    base = SmallMlp()
    fast = build_client(ScaffoldClient, copy.deepcopy(base), arena=True, lr=0.05)
    slow = build_client(ScaffoldClient, copy.deepcopy(base), arena=False, lr=0.05)
    assert isinstance(fast.optimizers["global"], FlatSGD)
    x, y = torch.randn(16, 8), torch.randint(0, 3, (16,))
    for client in (fast, slow):
        client.set_parameters(_packed_server_payload(client, 0.0), {"current_server_round": 1}, fitting_round=True)
        # pretend an earlier round left c_i != c: bump the local variates
        for v in client.client_control_variates:
            v.add_(0.01)
        client.set_parameters(_packed_server_payload(client, 0.02), {"current_server_round": 2}, fitting_round=True)
        for _ in range(3):
            client.train_step(x, y)
        client.update_after_train(3, {}, {})
    for a, b in zip(fast.model.parameters(), slow.model.parameters()):
        assert torch.allclose(a, b, atol=1e-6)
    for a, b in zip(fast.client_control_variates_updates, slow.client_control_variates_updates):
        assert torch.allclose(a, b, atol=1e-5)
    for a, b in zip(fast.client_control_variates, slow.client_control_variates):
        assert torch.allclose(a, b, atol=1e-5)
    # gradient correction really happened: first step moved weights by lr * (g + c - c_i), c - c_i = 0.01
    payload = fast.get_parameters({})
    assert len(payload) == len(fast.model.state_dict()) + len(list(fast.model.parameters()))


def test_scaffold_end_to_end_with_bn_and_frozen_layer() -> None:
    set_all_random_seeds(9)

    def model_fn():
        model = TinyNet()
        model.conv.requires_grad_(False)  # state_dict != trainable parameters
        return model

    clients = make_mixed_clients(ScaffoldClient, 2, model_fn=staticmethod(model_fn), momentum=0.0, lr=0.05)
    template = model_fn()
    torch.manual_seed(1234)
    strategy = Scaffold(
        initial_parameters=ndarrays_to_parameters([v.clone() for v in model_fn().state_dict().values()]),
        model=template, min_available_clients=2, on_fit_config_fn=fit_config_fn(), on_evaluate_config_fn=fit_config_fn(),
        fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
        evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn, learning_rate=1.0,
    )
    server = ScaffoldServer(SimpleClientManager(), {"n_server_rounds": 3}, strategy)
    history = run_simulation(server, clients, 3)
    losses = [l for _, l in history.losses_distributed]
    assert losses[-1] < losses[0], losses
    n_trainable = len([p for p in template.parameters() if p.requires_grad])
    assert len(strategy.server_control_variates) == n_trainable
    assert any(float(torch.as_tensor(v).abs().sum()) > 0 for v in strategy.server_control_variates)


def test_server_flat_step_matches_per_tensor_update() -> None:
    """Arena-shaped aggregates update the server's x and c with one op per block; same numbers as the per-tensor rule,
    including the integer BatchNorm counter, over consecutive rounds (the flat state persists between them)."""
    import copy

    from fl4health_b200.parallel.arena import TrainableRegionLayout, attach_arena

    torch.manual_seed(4)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.BatchNorm2d(4), torch.nn.Flatten(), torch.nn.Linear(4, 2))
    initial = [v.detach().clone() for v in net.state_dict().values()]
    make = lambda: Scaffold(initial_parameters=ndarrays_to_parameters([v.clone() for v in initial]), model=copy.deepcopy(net),  # noqa: E731
                            learning_rate=0.7, fraction_fit=0.5)
    flat_strategy, list_strategy = make(), make()
    arena = attach_arena(net)
    variates_layout = TrainableRegionLayout(arena)
    gen = torch.Generator().manual_seed(9)
    for round_index in range(3):
        aggregate_flat = torch.randn(arena.flat.shape, generator=gen)
        weights = NDArrays(arena.ndarrays(region=aggregate_flat), flat=aggregate_flat, layout=arena)
        counter_position = list(net.state_dict()).index("1.num_batches_tracked")
        weights[counter_position] = torch.tensor(10 * (round_index + 1))
        deltas_flat = torch.randn(arena.trainable_padded, generator=gen)
        deltas = variates_layout.ndarrays(region=deltas_flat)
        plain_weights, plain_deltas = NDArrays([w.clone() for w in weights]), NDArrays([d.clone() for d in deltas])
        new_x, new_c = flat_strategy.compute_updated_weights(weights), flat_strategy.compute_updated_control_variates(deltas)
        ref_x, ref_c = list_strategy.compute_updated_weights(plain_weights), list_strategy.compute_updated_control_variates(plain_deltas)
        assert getattr(new_x, "flat", None) is not None and getattr(ref_x, "flat", None) is None  # the two routes were taken
        for got, want in zip([*new_x, *new_c], [*ref_x, *ref_c]):
            assert got.dtype == want.dtype and torch.allclose(got.double(), want.double(), atol=1e-6)
        flat_strategy.server_model_weights, flat_strategy.server_control_variates = new_x, new_c
        list_strategy.server_model_weights, list_strategy.server_control_variates = ref_x, ref_c


def test_warm_start_hands_variates_to_clients_without_advancing_the_strategy() -> None:
    """Round 0 is a FIT whose answer is the regular packed payload (it used to be mistaken for the initial-parameters
    request); afterwards the clients hold warmed-up variates and the original weights, while the strategy's own running
    variates are still the initial ones -- the reference's behaviour, pinned by tests/differential/check_federations.py."""
    from fl4health_b200.client_managers.fixed_without_replacement_manager import FixedSamplingByFractionClientManager
    from fl4health_b200.simulation import register_clients

    set_all_random_seeds(11)
    clients = make_mixed_clients(ScaffoldClient, 2, model_fn=staticmethod(TinyNet), momentum=0.0, lr=0.05)
    template = TinyNet()
    initial = [v.clone() for v in template.state_dict().values()]
    strategy = Scaffold(
        initial_parameters=ndarrays_to_parameters([v.clone() for v in initial]), model=template, min_available_clients=2,
        on_fit_config_fn=fit_config_fn(), on_evaluate_config_fn=fit_config_fn(),
        fit_metrics_aggregation_fn=fit_metrics_aggregation_fn, evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn,
    )
    server = ScaffoldServer(FixedSamplingByFractionClientManager(), {"n_server_rounds": 1}, strategy, warm_start=True)
    register_clients(server, clients)
    packed = parameters_to_ndarrays(server._get_initial_parameters(server_round=0, timeout=None))
    weights, variates = strategy.parameter_packer.unpack_parameters(packed)
    assert all(torch.equal(torch.as_tensor(_np(w)), torch.as_tensor(_np(v))) for w, v in zip(weights, initial))
    assert any(float(torch.as_tensor(_np(v)).abs().sum()) > 0 for v in variates)  # warmed up
    assert all(float(torch.as_tensor(_np(v)).abs().sum()) == 0 for v in strategy.server_control_variates)  # not advanced
    assert all(client.client_control_variates is not None for client in clients)
