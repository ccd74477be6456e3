"""Self-attention for short sequences on the tcgen05 tensor cores (``csrc/attention_tc.cu``).

``packed_self_attention(qkv, key_mask, heads)`` takes the fused projection output ``[B, T, 3 * H]`` (``H = heads * 64``)
as is -- Q, K and V are addressed inside it by TMA, no split / permute copies -- and returns the context ``[B, T, H]``
ready for the output projection.  One CTA per (batch, head) for ``T <= 128``; backward recomputes the probabilities
from the saved log-sum-exp and writes the packed gradient ``[B, T, 3 * H]`` directly.
``packed_self_attention_reference`` is the ``scaled_dot_product_attention`` composition: fallback and numerics oracle.
"""

from __future__ import annotations

import ctypes
import math
import os

import torch
import torch.nn.functional as F_nn

from fl4health_b200.ops import _lib

HEAD_DIM = 64
MAX_SEQ = 128


def kernel_eligible(qkv: torch.Tensor, heads: int) -> bool:
    if not qkv.is_cuda or os.environ.get("FL4H_TC_ATTENTION", "1") == "0" or _lib.load() is None:
        return False
    if qkv.dtype != torch.bfloat16 or qkv.dim() != 3 or not qkv.is_contiguous():
        return False
    return qkv.shape[2] == 3 * heads * HEAD_DIM and 1 <= qkv.shape[1] <= MAX_SEQ


def packed_self_attention_reference(qkv: torch.Tensor, key_mask: torch.Tensor | None, heads: int) -> torch.Tensor:
    batch, seq, width = qkv.shape
    hidden = width // 3
    q, k, v = qkv.view(batch, seq, 3, heads, hidden // heads).permute(2, 0, 3, 1, 4)
    mask = key_mask[:, None, None, :].to(torch.bool) if key_mask is not None else None
    context = F_nn.scaled_dot_product_attention(q, k, v, attn_mask=mask)
    return context.transpose(1, 2).reshape(batch, seq, hidden)


class _PackedSelfAttention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, key_mask, heads):  # noqa: ANN001, ANN205
        lib = _lib.load(True)
        batch, seq, width = qkv.shape
        hidden = width // 3
        out = torch.empty(batch, seq, hidden, dtype=qkv.dtype, device=qkv.device)
        lse = torch.empty(batch, heads, seq, dtype=torch.float32, device=qkv.device)
        scale = 1.0 / math.sqrt(HEAD_DIM)
        err = lib.fl4h_attention_fwd(_lib.ptr(qkv), _lib.ptr(key_mask), _lib.ptr(out), _lib.ptr(lse), ctypes.c_int(batch),
                                     ctypes.c_int(seq), ctypes.c_int(heads), ctypes.c_float(scale), _lib.stream_ptr(qkv.device))
        _lib.check(err, "fl4h_attention_fwd")
        _lib.count_launches(1)
        ctx.save_for_backward(qkv, out, lse, key_mask if key_mask is not None else lse)
        ctx.conf = (heads, scale, key_mask is not None)
        return out

    @staticmethod
    def backward(ctx, grad_out):  # noqa: ANN001, ANN205
        qkv, out, lse, key_mask = ctx.saved_tensors
        heads, scale, masked = ctx.conf
        lib = _lib.load(True)
        batch, seq, _ = qkv.shape
        grad_out = grad_out.to(qkv.dtype).contiguous()
        dqkv = torch.empty_like(qkv)
        err = lib.fl4h_attention_bwd(_lib.ptr(qkv), _lib.ptr(key_mask if masked else None), _lib.ptr(out), _lib.ptr(grad_out),
                                     _lib.ptr(lse), _lib.ptr(dqkv), ctypes.c_int(batch), ctypes.c_int(seq), ctypes.c_int(heads),
                                     ctypes.c_float(scale), _lib.stream_ptr(qkv.device))
        _lib.check(err, "fl4h_attention_bwd")
        _lib.count_launches(1)
        return dqkv, None, None


def packed_self_attention(qkv: torch.Tensor, key_mask: torch.Tensor | None, heads: int) -> torch.Tensor:
    """``softmax(Q K^T / sqrt(d) + key padding) V`` for ``qkv = [B, T, 3 * heads * d]``; ``key_mask``: ``[B, T]``, non-zero =
    attend (uint8 for the kernel path)."""
    if kernel_eligible(qkv, heads) and (key_mask is None or (key_mask.dtype == torch.uint8 and key_mask.is_contiguous())):
        return _PackedSelfAttention.apply(qkv, key_mask, heads)
    return packed_self_attention_reference(qkv, key_mask, heads)
