"""Fused residual + dropout + LayerNorm (``ops/layer_norm.py`` / ``csrc/ln_fused.cu``) against plain PyTorch fp32."""

from __future__ import annotations

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(got: torch.Tensor, ref: torch.Tensor) -> float:
    return float((got.float() - ref.float()).abs().max() / ref.float().abs().max().clamp_min(1e-6))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("hidden", [256, 768, 1024])
@pytest.mark.parametrize("with_residual", [True, False])
def test_forward_backward_match_reference_without_dropout(dtype, hidden: int, with_residual: bool) -> None:  # noqa: ANN001
    from fl4health_b200.ops.layer_norm import add_dropout_layer_norm, add_dropout_layer_norm_reference, kernel_eligible

    torch.manual_seed(hidden)
    rows = 4 * 128 + 3  # not a multiple of the rows per CTA
    y = (torch.randn(rows, hidden, device="cuda") * 2 + 0.5).to(dtype).requires_grad_()
    res = torch.randn(rows, hidden, device="cuda").to(dtype).requires_grad_() if with_residual else None
    weight = (torch.rand(hidden, device="cuda") + 0.5).requires_grad_()
    bias = (torch.randn(hidden, device="cuda") * 0.1).requires_grad_()
    assert kernel_eligible(y, res, weight, bias)
    upstream = torch.randn(rows, hidden, device="cuda")
    out = add_dropout_layer_norm(y, res, weight, bias, 1e-12, 0.1, training=False)  # eval: dropout off
    assert out.dtype == dtype
    (out.float() * upstream).sum().backward()
    got = [out.detach(), y.grad, weight.grad, bias.grad] + ([res.grad] if with_residual else [])
    y32, w32, b32 = y.detach().float().requires_grad_(), weight.detach().clone().requires_grad_(), bias.detach().clone().requires_grad_()
    r32 = res.detach().float().requires_grad_() if with_residual else None
    ref_out = add_dropout_layer_norm_reference(y32, r32, w32, b32, 1e-12, 0.0, False)
    (ref_out * upstream).sum().backward()
    ref = [ref_out.detach(), y32.grad, w32.grad, b32.grad] + ([r32.grad] if with_residual else [])
    tol = 2e-2 if dtype == torch.bfloat16 else 2e-5
    for g, r in zip(got, ref):
        assert _rel(g, r) < tol


def test_dropout_mask_is_regenerated_in_backward_and_fresh_under_graph_replay() -> None:
    from fl4health_b200.ops.layer_norm import add_dropout_layer_norm

    torch.manual_seed(0)
    rows, hidden, p = 2048, 768, 0.25
    weight, bias = torch.ones(hidden, device="cuda", requires_grad=True), torch.zeros(hidden, device="cuda", requires_grad=True)
    # with residual == 0, gamma = 1, beta = 0 and y == 1 the kept units are exactly the positive outputs
    y = torch.ones(rows, hidden, device="cuda", requires_grad=True)
    out = add_dropout_layer_norm(y, None, weight, bias, 1e-5, p, training=True)
    kept = out.detach() > 0
    assert abs(float(kept.float().mean()) - (1 - p)) < 0.01
    out.sum().backward()  # d/dy of sum(LN(.)) is ~0 everywhere; use a non-degenerate upstream instead
    y.grad = None
    upstream = torch.randn(rows, hidden, device="cuda")
    out = add_dropout_layer_norm(y, None, weight, bias, 1e-5, p, training=True)
    kept = out.detach() > 0
    (out * upstream).sum().backward()
    assert bool(((y.grad != 0) <= kept).all())  # gradient flows only through units the forward kept (same mask, regenerated)
    assert float(((y.grad != 0) & kept).float().sum() / kept.float().sum()) > 0.99
    # replayed graphs draw new masks
    static_y = torch.ones(rows, hidden, device="cuda")
    graph = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.graph(graph), torch.no_grad():
        static_out = add_dropout_layer_norm(static_y, None, weight, bias, 1e-5, p, training=True)
    masks = []
    for _ in range(2):
        graph.replay()
        masks.append((static_out > 0).clone())
    assert not torch.equal(masks[0], masks[1]) and abs(float(masks[1].float().mean()) - (1 - p)) < 0.01


def test_bert_layer_uses_the_fused_epilogue_and_matches_the_stock_composition(monkeypatch) -> None:
    from fl4health_b200.models.bert import BertConfig, BertLayer

    torch.manual_seed(1)
    cfg = BertConfig(hidden_size=256, num_attention_heads=4, intermediate_size=512, hidden_dropout_prob=0.0)
    layer = BertLayer(cfg).cuda()
    x = torch.randn(4, 32, 256, device="cuda")
    out_kernel = layer(x, None)
    monkeypatch.setenv("FL4H_LN_KERNEL", "0")
    out_stock = layer(x, None)
    assert _rel(out_kernel, out_stock) < 1e-4
