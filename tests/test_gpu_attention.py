"""tcgen05 self-attention (``ops/attention.py`` / ``csrc/attention_tc.cu``) against a plain PyTorch fp32 reference."""

from __future__ import annotations

import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference(qkv: torch.Tensor, key_mask: torch.Tensor | None, heads: int) -> torch.Tensor:
    batch, seq, width = qkv.shape
    hidden = width // 3
    q, k, v = qkv.float().view(batch, seq, 3, heads, 64).permute(2, 0, 3, 1, 4)
    scores = q @ k.transpose(-1, -2) / math.sqrt(64)
    if key_mask is not None:
        scores = scores.masked_fill(~key_mask[:, None, None, :].bool(), float("-inf"))
    return (torch.softmax(scores, dim=-1) @ v).transpose(1, 2).reshape(batch, seq, hidden)


def _rel(got: torch.Tensor, ref: torch.Tensor) -> float:
    return float((got.float() - ref.float()).abs().max() / ref.float().abs().max().clamp_min(1e-6))


@pytest.mark.parametrize("seq", [128, 77, 16])
@pytest.mark.parametrize("masked", [False, True])
def test_forward_and_backward_match_reference(seq: int, masked: bool) -> None:
    from fl4health_b200.ops.attention import kernel_eligible, packed_self_attention

    torch.manual_seed(seq + masked)
    batch, heads = 3, 4
    qkv = (torch.randn(batch, seq, 3 * heads * 64, device="cuda") * 1.5).to(torch.bfloat16).requires_grad_()
    key_mask = None
    if masked:
        lengths = torch.tensor([seq, max(1, seq // 2), max(1, seq - 3)], device="cuda")
        key_mask = (torch.arange(seq, device="cuda")[None, :] < lengths[:, None]).to(torch.uint8)
    assert kernel_eligible(qkv, heads)
    upstream = torch.randn(batch, seq, heads * 64, device="cuda")
    out = packed_self_attention(qkv, key_mask, heads)
    assert out.dtype == torch.bfloat16 and out.shape == (batch, seq, heads * 64)
    (out.float() * upstream).sum().backward()
    ref_in = qkv.detach().float().requires_grad_()
    ref_out = _reference(ref_in, key_mask, heads)
    (ref_out * upstream).sum().backward()
    assert _rel(out, ref_out) < 2e-2
    got, want = qkv.grad.view(batch, seq, 3, heads * 64), ref_in.grad.view(batch, seq, 3, heads * 64)
    for which, name in enumerate("QKV"):
        assert _rel(got[:, :, which], want[:, :, which]) < 3e-2, name


def test_bert_layer_runs_on_the_attention_kernel(monkeypatch) -> None:
    from fl4health_b200.models.bert import BertConfig, BertLayer

    torch.manual_seed(2)
    cfg = BertConfig(hidden_size=256, num_attention_heads=4, intermediate_size=512, hidden_dropout_prob=0.0)
    layer = BertLayer(cfg).cuda()
    x = torch.randn(4, 48, 256, device="cuda")
    mask = (torch.arange(48, device="cuda")[None, :] < torch.tensor([48, 20, 33, 7], device="cuda")[:, None]).to(torch.uint8)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ours = layer(x.to(torch.bfloat16), mask)
        monkeypatch.setenv("FL4H_TC_ATTENTION", "0")
        stock = layer(x.to(torch.bfloat16), mask)
    valid = mask.bool()[:, :, None].expand_as(ours)
    assert _rel(ours[valid], stock[valid]) < 3e-2
