"""Master-weight (table) mode: per-tensor gradients + one multi-tensor optimizer launch must reproduce the flat-arena
optimizers exactly when the shadow is fp32, and train stably with a bf16 shadow."""

import copy

import pytest
import torch
from torch import nn

from fl4health_b200.engine.fused_optim import FlatAdamW, FlatSGD, translate_optimizer
from fl4health_b200.engine.options import EngineOptions
from fl4health_b200.parallel.arena import attach_arena
from tests.helpers import TinyNet, make_mixed_clients


def _model() -> nn.Module:
    torch.manual_seed(3)
    return nn.Sequential(nn.Conv2d(3, 5, 3, padding=1), nn.BatchNorm2d(5), nn.ReLU(), nn.Flatten(), nn.Linear(5 * 16, 7))


@pytest.mark.parametrize("kind", ["sgd", "adamw"])
@pytest.mark.parametrize("channels_last", [False, True])
def test_table_mode_matches_flat_mode(kind: str, channels_last: bool) -> None:
    base = _model()
    flat_model, table_model = copy.deepcopy(base), copy.deepcopy(base)
    flat_arena = attach_arena(flat_model, channels_last=channels_last)
    table_arena = attach_arena(table_model, channels_last=channels_last)
    table_arena.enable_compute_shadow(torch.float32)
    assert table_arena.grad is None and "0.weight" in table_arena.shadow_names and "1.weight" not in table_arena.shadow_names

    def make(arena):
        if kind == "sgd":
            return FlatSGD(arena, lr=0.1, momentum=0.9, weight_decay=1e-3)
        return FlatAdamW(arena, lr=1e-2, weight_decay=1e-2)

    opts = [make(flat_arena), make(table_arena)]
    assert opts[1].table_mode and not opts[0].table_mode
    anchors = [a.companion("anchor") for a in (flat_arena, table_arena)]
    for a, arena in zip(anchors, (flat_arena, table_arena)):
        a.copy_(arena.flat + 0.01)
    for opt, a in zip(opts, anchors):
        opt.set_drift_anchor(a, 0.5)
    gen = torch.Generator().manual_seed(0)
    for _ in range(4):
        x, y = torch.randn(6, 3, 4, 4, generator=gen), torch.randint(0, 7, (6,), generator=gen)
        for model, opt in zip((flat_model, table_model), opts):
            opt.zero_grad()
            nn.functional.cross_entropy(model(x), y).backward()
            opt.step()
        for entry in flat_arena.entries:  # (padding between entries is scratch: only real elements are compared)
            assert torch.allclose(flat_arena.view(entry.name), table_arena.view(entry.name), atol=1e-6), entry.name
        assert torch.equal(table_arena.view("0.weight", table_arena.shadow), table_arena.view("0.weight"))
    assert all(p.grad is None for p in table_model.parameters()) is False  # grads assigned by autograd
    opts[1].zero_grad()
    assert all(p.grad is None for p in table_model.parameters())


def test_translate_and_checkpoint_master() -> None:
    from fl4health_b200.checkpointing.checkpointer import materialize_module

    model = _model()
    arena = attach_arena(model)
    arena.enable_compute_shadow(torch.bfloat16)
    assert model[0].weight.dtype == torch.bfloat16 and model[1].weight.dtype == torch.float32
    opt = translate_optimizer(torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9), arena)
    assert isinstance(opt, FlatSGD) and opt.table_mode
    clone = materialize_module(model)
    assert clone[0].weight.dtype == torch.float32
    assert torch.equal(clone[0].weight.detach(), arena.view("0.weight"))
    # pulling new weights refreshes the compute shadow
    new = arena.ndarrays()
    arena.load_ndarrays([t + 1.0 if t.is_floating_point() else t for t in new])
    assert torch.allclose(model[0].weight.float(), arena.view("0.weight"), atol=2e-2)
    assert float(arena.view("0.weight").mean()) > 0.5


def test_bf16_master_weight_federation_cpu() -> None:
    from fl4health_b200.clients.fed_prox_client import FedProxClient
    from fl4health_b200.servers.base_server import FlServer  # noqa: F401
    from tests.test_fedprox import _run_fedprox  # type: ignore[attr-defined]

    engine = EngineOptions(amp_dtype=torch.bfloat16, master_weights=True)
    history, clients = _run_fedprox(FedProxClient, engine=engine)
    assert clients[0].optimizers["global"].table_mode
    losses = [v for _, v in history.losses_distributed]
    assert all(torch.isfinite(torch.tensor(losses))) and losses[-1] < losses[0] * 1.5


def test_flat_sgd_over_non_contiguous_ranges_matches_torch_with_dampening() -> None:
    """A parameter group that covers several separate arena ranges (FedRep head / base optimizers, sub-module
    optimizers) is stepped range by range with one hyper-parameter block: every range must take the first-step branch
    (momentum <- g), which only shows when dampening != 0."""
    torch.manual_seed(0)
    ours = nn.Sequential(nn.Linear(6, 8), nn.Linear(8, 8), nn.Linear(8, 8), nn.Linear(8, 3))
    stock = copy.deepcopy(ours)
    arena = attach_arena(ours)
    picked = lambda net: [*net[0].parameters(), *net[2].parameters()]  # noqa: E731 - layers 0 and 2: two separate ranges
    flat = FlatSGD(arena, picked(ours), lr=0.1, momentum=0.9, dampening=0.5)
    assert not flat.table_mode and len(flat._ranges[0]) >= 2
    reference = torch.optim.SGD(picked(stock), lr=0.1, momentum=0.9, dampening=0.5)
    gen = torch.Generator().manual_seed(1)
    for _ in range(3):
        x, y = torch.randn(5, 6, generator=gen), torch.randint(0, 3, (5,), generator=gen)
        for net, opt in ((ours, flat), (stock, reference)):
            opt.zero_grad()
            nn.functional.cross_entropy(net(x), y).backward()
            opt.step()
        for mine, theirs in zip(ours.parameters(), stock.parameters()):
            assert torch.allclose(mine, theirs, atol=1e-6)


@pytest.mark.parametrize("kind", ["sgd", "adamw"])
def test_fp32_table_gradients_get_the_fused_optimizer(kind: str) -> None:
    """``EngineOptions.table_grads`` (fp32, no flat gradient region, no shadow): the stock optimizer must still be
    translated to the one-launch pointer-table optimizer, and take the same steps as torch.optim -- with the FedProx
    anchor folded in (a stock optimizer would silently send drift-penalised clients down the autograd path)."""
    def net() -> nn.Module:  # (no convolution bias in front of BatchNorm: its true gradient is zero, Adam would amplify rounding noise)
        torch.manual_seed(3)
        return nn.Sequential(nn.Conv2d(3, 5, 3, padding=1, bias=False), nn.BatchNorm2d(5), nn.ReLU(), nn.Flatten(), nn.Linear(5 * 16, 7))

    ours, stock = net(), net()
    arena = attach_arena(ours)
    arena.use_table_gradients()
    assert arena.grad is None and arena.shadow is None and all(p.grad is None for p in ours.parameters())
    make = (lambda params: torch.optim.SGD(params, lr=0.1, momentum=0.9, weight_decay=1e-3)) if kind == "sgd" else \
        (lambda params: torch.optim.AdamW(params, lr=1e-2, weight_decay=1e-2))
    fused = translate_optimizer(make(ours.parameters()), arena)
    assert isinstance(fused, FlatSGD if kind == "sgd" else FlatAdamW) and fused.table_mode
    reference = make(stock.parameters())
    gen = torch.Generator().manual_seed(0)
    for _ in range(3):
        x, y = torch.randn(6, 3, 4, 4, generator=gen), torch.randint(0, 7, (6,), generator=gen)
        for net, opt in ((ours, fused), (stock, reference)):
            opt.zero_grad()
            nn.functional.cross_entropy(net(x), y).backward()
            opt.step()
        for mine, theirs in zip(ours.parameters(), stock.parameters()):
            assert torch.allclose(mine, theirs, atol=1e-5)
    assert all(p.grad is None for p in ours.parameters()) is False
    fused.zero_grad()
    assert all(p.grad is None for p in ours.parameters())
    frozen = attach_arena(_model(), with_grad=False)  # evaluation-only arenas keep whatever optimizer they are handed
    assert isinstance(translate_optimizer(torch.optim.SGD(frozen.module.parameters(), lr=0.1), frozen), torch.optim.SGD)
