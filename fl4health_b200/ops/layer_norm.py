"""``LayerNorm(residual + dropout(y))`` as one kernel forward and one backward (``csrc/ln_fused.cu``).

The block epilogue of every transformer sub-layer.  ``add_dropout_layer_norm_reference`` is the stock composition: the
CPU path, the fallback for unsupported shapes, and the numerics oracle for the GPU tests.  On the kernel path the
activations keep the input's dtype (bf16 in, bf16 out: no fp32 intermediates as under autocast), the dropout mask is
regenerated in the backward from a device-resident counter stream instead of being stored, and the stream advances on
the device, which makes the op CUDA-graph safe.
"""

from __future__ import annotations

import ctypes
import os

import torch
import torch.nn.functional as F_nn

from fl4health_b200.ops import _lib
from fl4health_b200.ops import flat as flat_ops

_DROPOUT_STREAMS: dict[int, torch.Tensor] = {}


def _dropout_stream(device: torch.device) -> torch.Tensor:
    index = device.index if device.index is not None else torch.cuda.current_device()
    if index not in _DROPOUT_STREAMS:
        _DROPOUT_STREAMS[index] = flat_ops.make_noise_state(device, torch.initial_seed() * 7919 + 17 * (index + 1))
    return _DROPOUT_STREAMS[index]


def kernel_eligible(y: torch.Tensor, residual: torch.Tensor | None, weight: torch.Tensor | None, bias: torch.Tensor | None) -> bool:
    if not y.is_cuda or os.environ.get("FL4H_LN_KERNEL", "1") == "0" or _lib.load() is None:
        return False
    hidden = y.shape[-1]
    if y.dtype not in (torch.bfloat16, torch.float32) or hidden % 256 != 0 or not 256 <= hidden <= 1024 or not y.is_contiguous():
        return False
    if residual is not None and (residual.shape != y.shape or residual.dtype != y.dtype or not residual.is_contiguous()):
        return False
    return weight is not None and bias is not None and weight.dtype == torch.float32 and bias.dtype == torch.float32


def add_dropout_layer_norm_reference(y: torch.Tensor, residual: torch.Tensor | None, weight: torch.Tensor | None,
                                     bias: torch.Tensor | None, eps: float, p: float, training: bool) -> torch.Tensor:
    pre = F_nn.dropout(y, p, training) if p > 0 else y
    if residual is not None:
        pre = residual + pre
    return F_nn.layer_norm(pre, (y.shape[-1],), weight, bias, eps)


class _AddDropoutLayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, residual, weight, bias, eps, p):  # noqa: ANN001, ANN205
        lib = _lib.load(True)
        hidden = y.shape[-1]
        rows = y.numel() // hidden
        out, pre = torch.empty_like(y), torch.empty_like(y)
        mean = torch.empty(rows, dtype=torch.float32, device=y.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=y.device)
        stream_state = _dropout_stream(y.device) if p > 0 else None
        used = torch.zeros(1, dtype=torch.int64, device=y.device) if p > 0 else None
        err = lib.fl4h_ln_fwd(_lib.ptr(y), _lib.ptr(residual), _lib.ptr(weight), _lib.ptr(bias), _lib.ptr(out), _lib.ptr(pre),
                              _lib.ptr(mean), _lib.ptr(rstd), ctypes.c_int64(rows), ctypes.c_int(hidden), ctypes.c_float(eps),
                              ctypes.c_float(p), _lib.ptr(stream_state), _lib.ptr(used),
                              ctypes.c_int(1 if y.dtype == torch.bfloat16 else 0), _lib.stream_ptr(y.device))
        _lib.check(err, "fl4h_ln_fwd")
        _lib.count_launches(2 if p > 0 else 1)
        ctx.save_for_backward(pre, mean, rstd, weight, used if used is not None else mean)
        ctx.conf = (p, residual is not None, stream_state)
        return out

    @staticmethod
    def backward(ctx, grad_out):  # noqa: ANN001, ANN205
        pre, mean, rstd, weight, used = ctx.saved_tensors
        p, has_residual, stream_state = ctx.conf
        lib = _lib.load(True)
        hidden = pre.shape[-1]
        rows = pre.numel() // hidden
        grad_out = grad_out.to(pre.dtype).contiguous()
        dpre = torch.empty_like(pre)
        dy = torch.empty_like(pre) if p > 0 else None
        dparams = torch.zeros(2, hidden, dtype=torch.float32, device=pre.device)
        err = lib.fl4h_ln_bwd(_lib.ptr(grad_out), _lib.ptr(pre), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(weight), _lib.ptr(dpre),
                              _lib.ptr(dy), _lib.ptr(dparams[0]), _lib.ptr(dparams[1]), ctypes.c_int64(rows), ctypes.c_int(hidden),
                              ctypes.c_float(p), _lib.ptr(stream_state), _lib.ptr(used if p > 0 else None),
                              ctypes.c_int(1 if pre.dtype == torch.bfloat16 else 0), _lib.stream_ptr(pre.device))
        _lib.check(err, "fl4h_ln_bwd")
        _lib.count_launches(1)
        return (dy if dy is not None else dpre), (dpre if has_residual else None), dparams[0], dparams[1], None, None


def add_dropout_layer_norm(y: torch.Tensor, residual: torch.Tensor | None, weight: torch.Tensor | None, bias: torch.Tensor | None,
                           eps: float = 1e-5, p: float = 0.0, training: bool = True) -> torch.Tensor:
    """``LayerNorm(residual + dropout(y, p))`` over the last dimension."""
    p = float(p) if training else 0.0
    if residual is not None and residual.dtype != y.dtype and y.is_cuda:
        residual = residual.to(y.dtype)
    if kernel_eligible(y, residual, weight, bias):
        return _AddDropoutLayerNorm.apply(y, residual, weight, bias, float(eps), p)
    return add_dropout_layer_norm_reference(y, residual, weight, bias, eps, p, training)
