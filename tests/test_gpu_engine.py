"""Engine paths on a real GPU: arena + fused optimizer + CUDA-graph step must reproduce eager PyTorch training."""

from pathlib import Path

import pytest
import torch

from fl4health_b200.engine.options import EngineOptions
from fl4health_b200.metrics.metric_aggregation import evaluate_metrics_aggregation_fn, fit_metrics_aggregation_fn
from fl4health_b200.servers.base_server import FlServer
from fl4health_b200.servers.client_manager import SimpleClientManager
from fl4health_b200.simulation import run_simulation
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg
from fl4health_b200.utils.random import set_all_random_seeds
from tests.helpers import TinyNet, fit_config_fn, make_clients

pytestmark = pytest.mark.gpu


def _run(engine: EngineOptions, rounds: int = 2, local_steps: int = 8, **client_kw):
    set_all_random_seeds(3)
    strategy = BasicFedAvg(min_fit_clients=2, min_evaluate_clients=2, min_available_clients=2,
                           on_fit_config_fn=fit_config_fn(local_steps), on_evaluate_config_fn=fit_config_fn(local_steps),
                           fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
                           evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn)
    server = FlServer(SimpleClientManager(), {"n_server_rounds": rounds}, strategy,
                      on_init_parameters_config_fn=fit_config_fn(local_steps))
    clients = make_clients(2, device="cuda", engine=engine, **client_kw)
    history = run_simulation(server, clients, num_rounds=rounds)
    state = {k: v.detach().float().cpu() for k, v in clients[0].model.state_dict().items()}
    return history, state, clients


@pytest.mark.parametrize("optimizer", ["sgd", "adamw"])
def test_arena_fused_matches_plain_eager(optimizer):
    torch.backends.cudnn.deterministic = True
    plain = EngineOptions(arena=False, fused_optimizer=False, cuda_graphs=False)
    fused = EngineOptions(arena=True, fused_optimizer=True, cuda_graphs=False)
    # Adam normalises by sqrt(v): parameters whose true gradient is zero (conv bias in front of BatchNorm) turn
    # rounding noise into +-lr steps, so the Adam comparison uses a BN-free model and looser tolerances.
    from fl4health_b200.models import Net

    model_fn, atol = (TinyNet, 2e-4) if optimizer == "sgd" else (Net, 2e-2)
    h0, s0, _ = _run(plain, model_fn=model_fn, optimizer=optimizer, lr=0.01 if optimizer == "sgd" else 1e-3)
    h1, s1, c1 = _run(fused, model_fn=model_fn, optimizer=optimizer, lr=0.01 if optimizer == "sgd" else 1e-3)
    from fl4health_b200.engine.fused_optim import _FlatOptimizer

    assert isinstance(c1[0].optimizers["global"], _FlatOptimizer)
    for key in s0:
        assert torch.allclose(s0[key], s1[key], atol=atol, rtol=1e-2), (key, (s0[key] - s1[key]).abs().max())
    assert abs(h0.losses_distributed[-1][1] - h1.losses_distributed[-1][1]) < (1e-3 if optimizer == "sgd" else 2e-2)


def test_cuda_graph_matches_eager():
    torch.backends.cudnn.deterministic = True
    eager = EngineOptions(cuda_graphs=False)
    graphed = EngineOptions(cuda_graphs=True, graph_warmup_steps=2)
    h0, s0, _ = _run(eager, rounds=3, model_fn=TinyNet)
    h1, s1, c1 = _run(graphed, rounds=3, model_fn=TinyNet)
    runner = c1[0]._train_runner
    assert runner is not None and runner.replays > 0, "train step was never replayed from a CUDA graph"
    for key in s0:
        assert torch.allclose(s0[key], s1[key], atol=2e-4, rtol=1e-3), (key, (s0[key] - s1[key]).abs().max())
    for (_, a), (_, b) in zip(h0.losses_distributed, h1.losses_distributed):
        assert abs(a - b) < 1e-3, (h0.losses_distributed, h1.losses_distributed)
    m0, m1 = h0.metrics_distributed["val - prediction - accuracy"], h1.metrics_distributed["val - prediction - accuracy"]
    assert all(abs(a[1] - b[1]) < 1e-6 for a, b in zip(m0, m1))


def test_bf16_channels_last_graph_trains():
    engine = EngineOptions(cuda_graphs=True, amp_dtype=torch.bfloat16, channels_last=True)
    history, _, clients = _run(engine, rounds=3, local_steps=12)
    losses = [l for _, l in history.losses_distributed]
    assert losses[-1] < losses[0], losses
    assert clients[0]._train_runner.replays > 0


def test_master_weights_graph_trains_like_autocast():
    """bf16 master-weight mode (table optimizer, bf16 parameter views) tracks plain bf16 autocast training."""
    auto = EngineOptions(cuda_graphs=True, amp_dtype=torch.bfloat16, channels_last=True)
    master = EngineOptions(cuda_graphs=True, amp_dtype=torch.bfloat16, channels_last=True, master_weights=True)
    h0, s0, _ = _run(auto, rounds=3, local_steps=12)
    h1, s1, c1 = _run(master, rounds=3, local_steps=12)
    opt = c1[0].optimizers["global"]
    assert opt.table_mode and c1[0]._train_runner.replays > 0
    losses = [l for _, l in h1.losses_distributed]
    assert losses[-1] < losses[0], losses
    for (_, a), (_, b) in zip(h0.losses_distributed, h1.losses_distributed):
        assert abs(a - b) < 0.15 * max(abs(a), 1.0), (h0.losses_distributed, h1.losses_distributed)
    # exchanged state is the fp32 master, and the compute shadow is its bf16 rounding
    arena = opt.arena
    name = next(iter(sorted(arena.shadow_names)))
    assert arena.view(name).dtype == torch.float32
    assert torch.allclose(arena.view(name, arena.shadow).float(), arena.view(name), atol=1e-2, rtol=1e-2)
