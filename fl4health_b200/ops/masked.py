"""Fused FedPM mask sampling (``csrc/masked_ops.cu``): ``out = Bernoulli(sigmoid(scores)) * frozen`` with the
straight-through gradient of the reference's ``BernoulliSample`` chained through the sigmoid."""

from __future__ import annotations

import ctypes
from typing import Any

import torch

from fl4health_b200.ops import _lib
from fl4health_b200.utils.functions import bernoulli_sample

_SEED_STATE: dict[int, torch.Tensor] = {}
_CALLS = 0


def _seed_state(device: torch.device) -> torch.Tensor:
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _SEED_STATE:
        _SEED_STATE[key] = torch.randint(0, 2**62, (1,), dtype=torch.int64).to(device)
    return _SEED_STATE[key]


def reseed(seed: int, device: torch.device) -> None:
    _seed_state(device).fill_(seed)


class _FusedMasked(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, scores: torch.Tensor, frozen: torch.Tensor) -> torch.Tensor:  # type: ignore[override]
        global _CALLS
        lib = _lib.load(True)
        s, w = scores.contiguous(), frozen.contiguous()
        out = torch.empty_like(w)
        _CALLS += 1
        err = lib.fl4h_masked_fwd(_lib.ptr(s), _lib.ptr(w), _lib.ptr(out), ctypes.c_void_p(0),
                                  _lib.ptr(_seed_state(s.device)), ctypes.c_uint64(_CALLS & 0xFFFF), ctypes.c_int64(s.numel()),
                                  ctypes.c_int(1), _lib.stream_ptr(s.device))
        _lib.check(err, "fl4h_masked_fwd")
        _lib.count_launches(2)
        ctx.save_for_backward(s, w)
        return out

    @staticmethod
    def backward(ctx: Any, grad_out: torch.Tensor) -> tuple:  # type: ignore[override]
        lib = _lib.load(True)
        s, w = ctx.saved_tensors
        grad_scores = torch.empty_like(s)
        err = lib.fl4h_masked_bwd(_lib.ptr(s), _lib.ptr(w), _lib.ptr(grad_out.contiguous()), _lib.ptr(grad_scores),
                                  ctypes.c_int64(s.numel()), _lib.stream_ptr(s.device))
        _lib.check(err, "fl4h_masked_bwd")
        _lib.count_launches(1)
        return grad_scores, None


def masked_parameter(scores: torch.Tensor, frozen: torch.Tensor) -> torch.Tensor:
    """Sample a binary mask from ``sigmoid(scores)`` and apply it to the frozen parameter."""
    if scores.is_cuda and scores.dtype == torch.float32 and frozen.dtype == torch.float32 and _lib.load() is not None:
        return _FusedMasked.apply(scores, frozen)
    return bernoulli_sample(torch.sigmoid(scores)) * frozen
