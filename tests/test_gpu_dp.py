"""DP-SGD on the GPU: device-resident noise stream under CUDA-graph replay, and the instance-level DP client running its
clip / noise / step inside captured graphs (one per Poisson batch size)."""

from __future__ import annotations

import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu


def test_noise_stream_advances_inside_a_replayed_graph() -> None:
    from fl4health_b200.ops import flat as F

    n = 1 << 20
    state = F.make_noise_state("cuda", seed=1234)
    buffer = torch.zeros(n, device="cuda")
    F.add_gaussian_state_(buffer, 2.0, state, 0.5)  # eager: (0 + 2 z) * 0.5 ~ N(0, 1)
    assert abs(float(buffer.mean())) < 0.01 and abs(float(buffer.std()) - 1.0) < 0.01
    assert state.tolist() == [1234, 1]
    graph, static = torch.cuda.CUDAGraph(), torch.zeros(n, device="cuda")
    torch.cuda.synchronize()
    with torch.cuda.graph(graph):
        static.zero_()
        F.add_gaussian_state_(static, 1.0, state, 1.0)
    draws = []
    for _ in range(3):
        graph.replay()
        draws.append(static.clone())
    torch.cuda.synchronize()
    assert state.tolist() == [1234, 4]
    for a, b in ((0, 1), (1, 2), (0, 2)):  # fresh, uncorrelated noise on every replay
        assert abs(float((draws[a] * draws[b]).mean())) < 0.01 and not torch.equal(draws[a], draws[b])
    assert all(abs(float(d.std()) - 1.0) < 0.01 for d in draws)
    same_stream = F.make_noise_state("cuda", seed=1234)
    same_stream[1] = 1
    again = torch.zeros(n, device="cuda")
    F.add_gaussian_state_(again, 1.0, same_stream, 1.0)
    assert torch.equal(again, draws[0])  # the stream is a pure function of (seed, position)


def test_ghost_clipping_matches_materialised_gradients_on_device() -> None:
    import copy

    from fl4health_b200.privacy.dp_engine import DPOptimizer, GradSampleModule

    torch.manual_seed(0)
    reference = nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.ReLU(), nn.AvgPool2d(4), nn.Conv2d(8, 32, 3, stride=2, padding=1),
                              nn.ReLU(), nn.Flatten(), nn.Linear(32 * 4 * 4, 64), nn.ReLU(), nn.Linear(64, 10)).cuda()
    ghost = copy.deepcopy(reference)
    x, y = torch.randn(16, 3, 32, 32, device="cuda") * 2, torch.randint(0, 10, (16,), device="cuda")
    stepped = {}
    for mode, model in (("hooks", reference), ("ghost", ghost)):
        wrapped = GradSampleModule(model, grad_sample_mode=mode)
        optimizer = DPOptimizer(torch.optim.SGD(model.parameters(), lr=0.1), noise_multiplier=0.0, max_grad_norm=0.5,
                                expected_batch_size=16, module=wrapped)
        optimizer.zero_grad()
        nn.functional.cross_entropy(wrapped(x), y).backward()
        optimizer.step()
        stepped[mode] = [p.detach().clone() for p in model.parameters()]
    for got, want in zip(stepped["ghost"], stepped["hooks"]):
        assert torch.allclose(got, want, rtol=1e-3, atol=1e-6)


def test_instance_level_dp_client_trains_inside_cuda_graphs() -> None:
    from fl4health_b200.clients.instance_level_dp_client import InstanceLevelDpClient
    from fl4health_b200.engine.options import EngineOptions
    from fl4health_b200.privacy.dp_engine import DPOptimizer, GradSampleModule
    from tests.helpers import TinyNet, make_mixed_clients

    engine = EngineOptions(arena=True, fused_optimizer=True, cuda_graphs=True, channels_last=False)
    (client,) = make_mixed_clients(InstanceLevelDpClient, 1, device="cuda", engine=engine, model_fn=staticmethod(TinyNet),
                                   momentum=0.0, lr=0.05)
    config = {"current_server_round": 1, "local_steps": 40, "batch_size": 32, "clipping_bound": 1.0, "noise_multiplier": 0.5}
    client.setup_client(config)
    assert isinstance(client.model, GradSampleModule) and client.model.grad_sample_mode == "ghost"
    assert isinstance(client.optimizers["global"], DPOptimizer)
    before = torch.cat([p.detach().flatten().clone() for p in client.model.parameters()])
    parameters = client.get_parameters({**config, "current_server_round": 2})
    _, _, metrics = client.fit(parameters, config)
    runner = client._executor.latest_train_runner
    assert runner is not None and runner.replays > 0 and len(runner._graphs) >= 1, (runner.replays, runner.eager_steps)
    after = torch.cat([p.detach().flatten() for p in client.model.parameters()])
    assert torch.isfinite(after).all() and float((after - before).abs().max()) > 0
    state = client.optimizers["global"]._noise_state
    assert state is not None and int(state[1]) == 40  # one draw per step, replayed steps included
