"""MOON client + contrastive losses (oracles follow the reference's tests/losses and tests/clients/test_moon_client.py)."""

import math

import pytest
import torch
from torch import nn

from fl4health_b200.clients.moon_client import MoonClient
from fl4health_b200.losses.contrastive_loss import MoonContrastiveLoss, NtXentLoss
from fl4health_b200.losses.cosine_similarity_loss import CosineSimilarityLoss
from fl4health_b200.losses.perfcl_loss import PerFclLoss
from fl4health_b200.metrics.metric_aggregation import evaluate_metrics_aggregation_fn, fit_metrics_aggregation_fn
from fl4health_b200.model_bases.moon_base import MoonModel
from fl4health_b200.ops.contrastive import moon_contrastive, moon_contrastive_reference
from fl4health_b200.servers.base_server import FlServer
from fl4health_b200.servers.client_manager import SimpleClientManager
from fl4health_b200.simulation import run_simulation
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg
from fl4health_b200.utils.random import set_all_random_seeds
from tests.helpers import fit_config_fn, make_mixed_clients

CPU = torch.device("cpu")


def test_moon_contrastive_closed_form() -> None:
    loss_fn = MoonContrastiveLoss(CPU, temperature=0.5)
    z = torch.tensor([[1.0, 0.0], [0.0, 2.0]])
    pos = z.clone().unsqueeze(0)  # cos = 1
    neg = (-z).unsqueeze(0)  # cos = -1
    # logits = [2, -2] -> CE = log(1 + e^-4)
    assert loss_fn(z, pos, neg).item() == pytest.approx(math.log(1 + math.exp(-4.0)), rel=1e-5)
    orth = torch.tensor([[0.0, 1.0], [3.0, 0.0]]).unsqueeze(0)  # cos = 0
    assert loss_fn(z, orth, orth).item() == pytest.approx(math.log(2.0), rel=1e-5)
    with pytest.raises(AssertionError):
        loss_fn(z, torch.cat([pos, pos]), neg)


def test_moon_contrastive_gradients_match_autograd_reference() -> None:
    torch.manual_seed(0)
    z = torch.randn(6, 10, requires_grad=True)
    pos, neg = torch.randn(6, 10, requires_grad=True), torch.randn(3, 6, 10, requires_grad=True)
    ref = moon_contrastive_reference(z, pos, neg, 0.7)
    ref.backward()
    z2, p2, n2 = (t.detach().clone().requires_grad_() for t in (z, pos, neg))
    out = moon_contrastive(z2, p2, n2, 0.7)
    out.backward()
    assert out.item() == pytest.approx(ref.item(), rel=1e-6)
    for a, b in ((z, z2), (pos, p2), (neg, n2)):
        assert torch.allclose(a.grad, b.grad, atol=1e-6)


def test_ntxent_cosine_perfcl() -> None:
    f = torch.tensor([[1.0, 0.0], [0.0, 1.0]])
    loss = NtXentLoss(CPU, temperature=1.0)(f, f)
    # every row: positive sim 1, one orthogonal pair twice (0), diagonal zeroed -> denom = e^1 + 2e^0 + e^0
    assert loss.item() == pytest.approx(-1.0 + math.log(math.e + 3.0), rel=1e-5)
    assert CosineSimilarityLoss(CPU)(f, -f).item() == pytest.approx(1.0)
    g, l = PerFclLoss(CPU)(f, f, f, -f, f)
    assert g.item() == pytest.approx(math.log(1 + math.exp(-4.0)), rel=1e-5)  # pos cos 1, neg cos -1
    assert l.item() == pytest.approx(math.log(2.0), rel=1e-5)  # pos == neg


class _Base(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.conv = nn.Conv2d(3, 4, 3, padding=1)

    def forward(self, x):
        return torch.nn.functional.adaptive_avg_pool2d(torch.relu(self.conv(x)), 4)


class _Head(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.fc = nn.Linear(64, 10)

    def forward(self, x):
        return self.fc(torch.flatten(x, 1))


def test_moon_client_end_to_end() -> None:
    set_all_random_seeds(31)
    clients = make_mixed_clients(MoonClient, 2, model_fn=staticmethod(lambda: MoonModel(_Base(), _Head())))
    for c in clients:
        c.len_old_models_buffer = 2
    common = dict(min_fit_clients=2, min_evaluate_clients=2, min_available_clients=2, on_fit_config_fn=fit_config_fn(),
                  on_evaluate_config_fn=fit_config_fn(), fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
                  evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn)
    server = FlServer(SimpleClientManager(), {"n_server_rounds": 3}, BasicFedAvg(**common),
                      on_init_parameters_config_fn=fit_config_fn())
    history = run_simulation(server, clients, 3)
    assert len(history.losses_distributed) == 3
    c0 = clients[0]
    assert len(c0.old_models_list) == 2 and c0.global_model is not None
    assert all(not p.requires_grad for m in c0.old_models_list for p in m.parameters())
    # from round 2 on the contrastive term is active and reported
    x, y = torch.randn(4, 3, 32, 32), torch.randint(0, 10, (4,))
    c0.model.train()
    losses, _ = c0.train_step(x, y)
    assert "contrastive_loss" in losses.additional_losses and losses.additional_losses["contrastive_loss"].item() > 0
