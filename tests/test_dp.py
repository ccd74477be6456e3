"""Differential privacy: per-sample gradient engine vs a per-example loop, flat clipping semantics, client-level
clipping oracles (reference: tests/clients/test_clipping_client.py, tests/strategies/test_client_dp_fedavgm.py), and
end-to-end instance-level / client-level DP federations."""

from pathlib import Path

import numpy as np
import pytest
import torch
from torch import nn

from fl4health_b200.client_managers.fixed_without_replacement_manager import FixedSamplingByFractionClientManager
from fl4health_b200.client_managers.poisson_sampling_manager import PoissonSamplingClientManager
from fl4health_b200.clients.clipping_client import NumpyClippingClient
from fl4health_b200.clients.instance_level_dp_client import InstanceLevelDpClient
from fl4health_b200.clients.scaffold_client import DPScaffoldClient
from fl4health_b200.common.typing import NDArrays, Parameters, to_numpy
from fl4health_b200.metrics import Accuracy
from fl4health_b200.metrics.metric_aggregation import evaluate_metrics_aggregation_fn, fit_metrics_aggregation_fn
from fl4health_b200.privacy.dp_engine import DPOptimizer, GradSampleModule, ModuleValidator, PrivacyEngine
from fl4health_b200.servers.client_level_dp_fed_avg_server import ClientLevelDPFedAvgServer
from fl4health_b200.servers.instance_level_dp_server import InstanceLevelDpServer
from fl4health_b200.simulation import run_simulation
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg
from fl4health_b200.strategies.client_dp_fedavgm import ClientLevelDPFedAvgM
from fl4health_b200.utils.random import set_all_random_seeds
from tests.helpers import TinyNet, make_mixed_clients


class DpNet(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.emb = nn.Embedding(7, 4)
        self.conv = nn.Conv2d(3, 4, 3, padding=1, stride=2)
        self.gn = nn.GroupNorm(2, 4)
        self.conv1d = nn.Conv1d(4, 4, 3, padding=1)
        self.ln = nn.LayerNorm(8)
        self.fc = nn.Linear(8, 5)

    def forward(self, x: torch.Tensor, tokens: torch.Tensor) -> torch.Tensor:
        h = torch.relu(self.gn(self.conv(x)))  # [B,4,4,4]
        h = self.conv1d(h.mean(dim=3))  # [B,4,4]
        h = h.mean(dim=2)  # [B,4]
        e = self.emb(tokens).mean(dim=1)  # [B,4]
        return self.fc(self.ln(torch.cat([h, e], dim=1)))


def test_per_sample_gradients_match_per_example_loop() -> None:
    torch.manual_seed(0)
    model = DpNet()
    wrapped = GradSampleModule(model)
    x, tok, y = torch.randn(6, 3, 8, 8), torch.randint(0, 7, (6, 3)), torch.randint(0, 5, (6,))
    wrapped.train()
    nn.functional.cross_entropy(wrapped(x, tok), y).backward()
    per_sample = {name: p.grad_sample.clone() for name, p in model.named_parameters()}
    for i in range(6):
        model.zero_grad()
        wrapped.hooks_enabled = False
        nn.functional.cross_entropy(model(x[i : i + 1], tok[i : i + 1]), y[i : i + 1]).backward()
        for name, p in model.named_parameters():
            assert torch.allclose(per_sample[name][i], p.grad, atol=1e-5), (name, i)


def test_dp_optimizer_flat_clipping_without_noise() -> None:
    torch.manual_seed(1)
    model = nn.Sequential(nn.Flatten(), nn.Linear(12, 3))
    wrapped = GradSampleModule(model)
    inner = torch.optim.SGD(model.parameters(), lr=1.0)
    opt = DPOptimizer(inner, noise_multiplier=0.0, max_grad_norm=0.5, expected_batch_size=4, module=wrapped)
    x, y = torch.randn(4, 3, 2, 2), torch.randint(0, 3, (4,))
    before = [p.detach().clone() for p in model.parameters()]
    opt.zero_grad()
    nn.functional.cross_entropy(wrapped(x), y).backward()
    samples = [p.grad_sample.clone() for p in model.parameters()]
    opt.step()
    norms = torch.stack([s.reshape(4, -1).norm(dim=1) for s in samples], 1).norm(dim=1)
    factor = (0.5 / (norms + 1e-6)).clamp(max=1.0)
    for p, b, s in zip(model.parameters(), before, samples):
        expected = b - torch.einsum("i,i...", factor, s) / 4
        assert torch.allclose(p.detach(), expected, atol=1e-6)
    assert float((factor * norms).max()) <= 0.5 + 1e-5


def test_module_validator_replaces_batchnorm() -> None:
    model = TinyNet()
    assert not ModuleValidator.is_valid(model)
    fixed = ModuleValidator.fix(model)
    assert isinstance(fixed.bn, nn.GroupNorm) and ModuleValidator.is_valid(fixed)
    wrapped, opt, loader = PrivacyEngine().make_private(
        module=fixed, optimizer=torch.optim.SGD(fixed.parameters(), lr=0.1),
        data_loader=type("L", (), {"dataset": list(range(100)), "batch_size": 10})(), noise_multiplier=1.0,
        max_grad_norm=1.0, poisson_sampling=False)
    assert all(k.startswith("_module.") for k in wrapped.state_dict())


def test_clipping_client_oracles() -> None:
    client = NumpyClippingClient(Path(""), [Accuracy("accuracy")], torch.device("cpu"))
    client.adaptive_clipping, client.clipping_bound = True, 1.0
    client.initial_weights = NDArrays([2.0 * np.ones((2, 3, 3)) for _ in range(4)])
    update, bit = client.compute_weight_update_and_clip(NDArrays([4.0 * np.ones((2, 3, 3)) for _ in range(4)]))
    assert bit == 0.0 and to_numpy(update[0])[0, 0, 0] == pytest.approx(0.11785, abs=1e-4)
    client.clipping_bound = 9.0
    update, bit = client.compute_weight_update_and_clip(NDArrays([3.0 * np.ones((2, 3, 3)) for _ in range(4)]))
    assert bit == 1.0 and to_numpy(update[0])[0, 0, 0] == pytest.approx(1.0, abs=1e-4)


def test_client_dp_fedavgm_oracles() -> None:
    strategy = ClientLevelDPFedAvgM(initial_parameters=Parameters([], ""), adaptive_clipping=True, server_learning_rate=0.5,
                                    clipping_learning_rate=0.5, weight_noise_multiplier=2.0, clipping_noise_multiplier=5.0)
    assert strategy.modify_noise_multiplier() == pytest.approx(2.0412, abs=1e-4)
    np.random.seed(42)
    updates = NDArrays([np.random.rand(2, 3) for _ in range(4)])
    strategy.calculate_update_with_momentum(updates)
    strategy.calculate_update_with_momentum(updates)
    for m, u in zip(strategy.m_t, updates):
        assert np.allclose(to_numpy(m), (1.0 + strategy.beta) * u)
    # clipping-bound update rule with a noiseless bit mean: C <- C exp(-lr (b - gamma))
    strategy.clipping_bound = 0.1
    strategy._update_clipping_bound_with_noised_bits(0.8)
    assert strategy.clipping_bound == pytest.approx(0.1 * np.exp(-0.5 * (0.8 - 0.5)))


def _dp_config(r):
    return {"current_server_round": r, "local_steps": 3, "batch_size": 32, "clipping_bound": 1.0, "noise_multiplier": 0.5,
            "adaptive_clipping": True}


def test_instance_level_dp_end_to_end() -> None:
    set_all_random_seeds(61)
    clients = make_mixed_clients(InstanceLevelDpClient, 2, model_fn=staticmethod(TinyNet), momentum=0.0, lr=0.05)
    strategy = BasicFedAvg(fraction_fit=1.0, fraction_evaluate=1.0, min_fit_clients=2, min_evaluate_clients=2,
                           min_available_clients=2, on_fit_config_fn=_dp_config, on_evaluate_config_fn=_dp_config,
                           fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
                           evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn)
    server = InstanceLevelDpServer(PoissonSamplingClientManager(), {"n_server_rounds": 2}, noise_multiplier=0.5, batch_size=32,
                                   num_server_rounds=2, strategy=strategy, local_steps=3, on_init_parameters_config_fn=_dp_config)
    history = run_simulation(server, clients, 2)
    assert len(history.losses_distributed) == 2
    assert isinstance(clients[0].model, GradSampleModule) and isinstance(clients[0].optimizers["global"], DPOptimizer)
    assert isinstance(clients[0].model._module.bn, nn.GroupNorm)
    assert server.accountant.get_epsilon(2, 1e-3) > 0


def test_dp_scaffold_client_setup() -> None:
    set_all_random_seeds(62)
    client = make_mixed_clients(DPScaffoldClient, 1, model_fn=staticmethod(TinyNet), momentum=0.0, lr=0.05)[0]
    client.setup_client(_dp_config(1))
    assert isinstance(client.model, GradSampleModule) and isinstance(client.optimizers["global"], DPOptimizer)
    assert client.learning_rate == 0.05 and hasattr(client, "client_control_variates")


@pytest.mark.parametrize("manager_cls", [PoissonSamplingClientManager, FixedSamplingByFractionClientManager])
def test_client_level_dp_end_to_end(manager_cls) -> None:
    set_all_random_seeds(63)
    clients = make_mixed_clients(NumpyClippingClient, 3, model_fn=staticmethod(TinyNet))
    strategy = ClientLevelDPFedAvgM(fraction_fit=1.0, fraction_evaluate=1.0, min_available_clients=3,
                                    on_fit_config_fn=_dp_config, on_evaluate_config_fn=_dp_config,
                                    fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
                                    evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn, adaptive_clipping=True,
                                    initial_clipping_bound=0.5, weight_noise_multiplier=0.1, clipping_noise_multiplier=5.0,
                                    weighted_aggregation=manager_cls is PoissonSamplingClientManager)
    server = ClientLevelDPFedAvgServer(manager_cls(), {"n_server_rounds": 2}, strategy, server_noise_multiplier=0.1,
                                       num_server_rounds=2, on_init_parameters_config_fn=lambda r: _dp_config(0))
    history = run_simulation(server, clients, 2)
    assert len(history.losses_distributed) == 2
    assert strategy.clipping_bound != 0.5  # adapted
    s0, s1 = clients[0].model.state_dict(), clients[1].model.state_dict()
    assert all(torch.allclose(s0[k].float(), s1[k].float()) for k in s0)


class _GramNet(nn.Module):
    """Layers on both sides of the ghost-norm decision: a convolution with few output positions and a per-token Linear
    with few tokens take the Gram identity, the first convolution (many positions, tiny filter) forms its per-sample
    gradients, the classifier sees one token per sample."""

    def __init__(self) -> None:
        super().__init__()
        self.stem = nn.Conv2d(3, 8, 3, padding=1)            # 64 positions, 27 x 8 filter: direct
        self.deep = nn.Conv2d(8, 32, 3, stride=2, padding=1)   # 4 positions, 72 x 32 filter: Gram
        self.token_mlp = nn.Linear(32, 48)                      # 4 tokens, 32 x 48: Gram
        self.head = nn.Linear(48, 5)                            # 1 token: norm product

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        h = torch.relu(self.stem(x))
        h = torch.relu(self.deep(nn.functional.avg_pool2d(h, 2)))  # [B, 32, 2, 2]
        tokens = torch.relu(self.token_mlp(h.flatten(2).transpose(1, 2)))  # [B, 4, 48]
        return self.head(tokens.mean(dim=1))


@pytest.mark.parametrize("net", [DpNet, _GramNet])
def test_ghost_book_keeping_matches_materialised_per_sample_gradients(net) -> None:  # noqa: ANN001
    """``grad_sample_mode="ghost"`` (norms via Gram identities, clipped sum as one GEMM per layer, no [B, *shape]
    tensors for the factored layers) takes the same DP-SGD step as the Opacus-style materialising hooks."""
    import copy

    from fl4health_b200.privacy import dp_engine

    torch.manual_seed(5)
    reference_model = net()
    ghost_model = copy.deepcopy(reference_model)
    batch = 6
    if net is DpNet:
        inputs = (torch.randn(batch, 3, 8, 8), torch.randint(0, 7, (batch, 3)))
    else:
        inputs = (torch.randn(batch, 3, 8, 8) * 3,)
    labels = torch.randint(0, 5, (batch,))
    taken: dict[str, list] = {}
    for mode, model in (("hooks", reference_model), ("ghost", ghost_model)):
        wrapped = GradSampleModule(model, grad_sample_mode=mode)
        optimizer = DPOptimizer(torch.optim.SGD(model.parameters(), lr=0.5), noise_multiplier=0.0, max_grad_norm=0.7,
                                expected_batch_size=batch, module=wrapped)
        for _ in range(2):  # two steps: the book is emptied and refilled
            optimizer.zero_grad()
            nn.functional.cross_entropy(wrapped(*inputs), labels).backward()
            if mode == "ghost":
                assert all(getattr(p, "grad_sample", None) is None for p in model.parameters())
                factored = [kept.weight is not None for kept in wrapped.deferred]
                assert any(factored) and (net is DpNet or not all(factored))
                norms = wrapped.sq_norms.sqrt()
            else:
                norms = torch.stack([p.grad_sample.reshape(batch, -1).norm(dim=1) for p in model.parameters()], 1).norm(dim=1)
            taken.setdefault(mode + "_norms", []).append(norms.clone())
            optimizer.step()
        taken[mode] = [p.detach().clone() for p in model.parameters()]
    for got, want in zip(taken["ghost_norms"], taken["hooks_norms"]):
        assert torch.allclose(got, want, rtol=1e-4, atol=1e-6)
    assert float(taken["hooks_norms"][0].max()) > 0.7  # clipping was active
    for got, want in zip(taken["ghost"], taken["hooks"]):
        assert torch.allclose(got, want, rtol=1e-4, atol=1e-6)
    assert dp_engine._gram_is_cheaper(4, 72, 32) and not dp_engine._gram_is_cheaper(64, 27, 8)
