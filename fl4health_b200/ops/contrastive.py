"""Fused MOON-style contrastive loss (``csrc/moon_loss.cu``) as a differentiable op + its PyTorch reference."""

from __future__ import annotations

import ctypes
from typing import Any

import torch

from fl4health_b200.ops import _lib

MAX_NEGATIVES = 15


def moon_contrastive_reference(z: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor, temperature: float) -> torch.Tensor:
    """z [B,F], pos [B,F], neg [N,B,F] -> scalar CE over cosine logits with the positive in slot 0."""
    cos = torch.nn.functional.cosine_similarity
    logits = torch.cat([cos(z, pos, dim=-1).unsqueeze(1), cos(z.unsqueeze(0), neg, dim=-1).t()], dim=1) / temperature
    labels = torch.zeros(z.shape[0], dtype=torch.long, device=z.device)
    return torch.nn.functional.cross_entropy(logits, labels)


class _FusedMoon(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, z: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor, temperature: float) -> torch.Tensor:  # type: ignore[override]
        lib = _lib.load(True)
        batch, feat = z.shape
        n_neg = neg.shape[0]
        zc, pc, nc = z.contiguous().float(), pos.contiguous().float(), neg.contiguous().float()
        loss = torch.zeros((), dtype=torch.float32, device=z.device)
        probs = torch.empty(batch, 1 + n_neg, dtype=torch.float32, device=z.device)
        cosines = torch.empty_like(probs)
        norms = torch.empty(batch, 2 + n_neg, dtype=torch.float32, device=z.device)
        err = lib.fl4h_moon_fwd(_lib.ptr(zc), _lib.ptr(pc), _lib.ptr(nc), ctypes.c_int(batch), ctypes.c_int(feat),
                                ctypes.c_int(n_neg), ctypes.c_float(temperature), _lib.ptr(loss), _lib.ptr(probs),
                                _lib.ptr(cosines), _lib.ptr(norms), _lib.stream_ptr(z.device))
        _lib.check(err, "fl4h_moon_fwd")
        _lib.count_launches(1)
        ctx.save_for_backward(zc, pc, nc, probs, cosines, norms)
        ctx.temperature = temperature
        ctx.dtypes = (z.dtype, pos.dtype, neg.dtype)
        return loss

    @staticmethod
    def backward(ctx: Any, grad_out: torch.Tensor) -> tuple:  # type: ignore[override]
        lib = _lib.load(True)
        zc, pc, nc, probs, cosines, norms = ctx.saved_tensors
        batch, feat = zc.shape
        need_z, need_p, need_n = ctx.needs_input_grad[:3]
        gz = torch.empty_like(zc) if need_z else None
        gp = torch.empty_like(pc) if need_p else None
        gn = torch.empty_like(nc) if need_n else None
        go = grad_out.contiguous().float().reshape(1)
        err = lib.fl4h_moon_bwd(_lib.ptr(zc), _lib.ptr(pc), _lib.ptr(nc), _lib.ptr(probs), _lib.ptr(cosines),
                                _lib.ptr(norms), _lib.ptr(go), ctypes.c_int(batch), ctypes.c_int(feat),
                                ctypes.c_int(nc.shape[0]), ctypes.c_float(ctx.temperature), _lib.ptr(gz), _lib.ptr(gp),
                                _lib.ptr(gn), _lib.stream_ptr(zc.device))
        _lib.check(err, "fl4h_moon_bwd")
        _lib.count_launches(1)
        dz, dp, dn = ctx.dtypes
        return (gz.to(dz) if gz is not None else None, gp.to(dp) if gp is not None else None,
                gn.to(dn) if gn is not None else None, None)


def moon_contrastive(z: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor, temperature: float) -> torch.Tensor:
    if z.is_cuda and z.dim() == 2 and neg.shape[0] <= MAX_NEGATIVES and _lib.load() is not None:
        return _FusedMoon.apply(z, pos, neg, float(temperature))
    return moon_contrastive_reference(z, pos, neg, temperature)
