"""Fused BatchNorm2d (+ residual add) (+ ReLU) for channels-last activations (kernels: ``csrc/bn_act.cu``).

``batch_norm_act(x, bn_module_state..., residual, relu)`` is an autograd function whose training forward is ONE
cooperative launch (reduce -> grid barrier -> apply) and whose backward is one; the stock composition (``F.batch_norm`` + ``add`` + ``relu`` and their
backward, plus the running-stat and ``num_batches_tracked`` updates) is ~10 launches per layer and step, which is what a
CIFAR-scale ResNet client is bound by.  ``batch_norm_act_reference`` is that stock composition: the CPU path, the
fallback for unsupported layouts, and the numerics oracle for the GPU tests.
"""

from __future__ import annotations

import ctypes
import os

import torch
import torch.nn.functional as F_nn

from fl4health_b200.ops import _lib

_WORKSPACES: dict[tuple[int, int], tuple[torch.Tensor, torch.Tensor]] = {}
_ACC_STRIDE = 4096  # floats; must match kAccStride in bn_act.cu (supports C <= 2048)


def _allow_fused() -> int:
    """One cooperative kernel per direction (default) vs the two-kernel chain (``FL4H_BN_FUSED=0``, for A/B runs)."""
    return 0 if os.environ.get("FL4H_BN_FUSED", "1") == "0" else 1


def _workspace(device: torch.device, channels: int) -> tuple[torch.Tensor, torch.Tensor]:
    """Persistent zero-initialised accumulators + election counter (self-resetting: every kernel leaves them zero)."""
    assert 2 * channels <= _ACC_STRIDE
    # one workspace per (device, stream): kernels of one stream serialise, so they can share accumulators
    key = (device.index if device.index is not None else torch.cuda.current_device(), torch.cuda.current_stream(device).cuda_stream)
    ws = _WORKSPACES.get(key)
    if ws is None:
        # [2 x stride] double-buffered accumulators of the fused kernels + [stride] for the two-kernel fallback;
        # counter[0] = last-CTA election, counter[1] = accumulator parity
        ws = (torch.zeros(3 * _ACC_STRIDE, dtype=torch.float32, device=device), torch.zeros(4, dtype=torch.int32, device=device))
        _WORKSPACES[key] = ws
    return ws


def _is_channels_last_dense(t: torch.Tensor) -> bool:
    return t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last)


def _kernels_enabled() -> bool:
    """``FL4H_BN_KERNEL=0`` sends every BatchNorm through the stock-op reference path (A/B runs, numerics tests)."""
    return os.environ.get("FL4H_BN_KERNEL", "1") != "0"


def kernel_eligible(x: torch.Tensor, residual: torch.Tensor | None, momentum: float | None, training: bool,
                    running_mean: torch.Tensor | None) -> bool:
    if not x.is_cuda or not _kernels_enabled() or _lib.load() is None:
        return False
    if x.dtype not in (torch.bfloat16, torch.float32) or not _is_channels_last_dense(x):
        return False
    n, c, h, w = x.shape
    if c % 8 != 0 or c // 8 > 256:
        return False
    if residual is not None and (residual.shape != x.shape or residual.dtype != x.dtype or not _is_channels_last_dense(residual)):
        return False
    if training and running_mean is not None and momentum is None:
        return False  # cumulative moving average: stock path
    if not training and running_mean is None:
        return False
    if running_mean is not None and running_mean.dtype != torch.float32:
        return False  # the kernels keep statistics in fp32 (a model cast wholesale to bf16 takes the stock path)
    if training and os.environ.get("FL4H_DETERMINISTIC", "0") == "1":
        # the grid-wide kernels accumulate with fp32 RED atomics (order varies run to run): deterministic mode takes the
        # (deterministic) ATen path for training statistics
        return False
    return True


def presums_eligible(x_like: torch.Tensor, channels: int, residual: torch.Tensor | None, momentum: float | None,
                     training: bool, running_mean: torch.Tensor | None) -> bool:
    """True when ``batch_norm_act(..., presums=...)`` will take the kernel path for a tensor shaped like ``x_like``
    with ``channels`` channels (decided BEFORE the convolution runs, so its epilogue knows whether to reduce)."""
    if not training or not x_like.is_cuda or not _kernels_enabled() or _lib.load() is None:
        return False
    if x_like.dtype not in (torch.bfloat16, torch.float32):
        return False
    if channels % 8 != 0 or channels // 8 > 256 or 256 % (channels // 8) != 0:
        return False
    if residual is not None and (residual.dtype != x_like.dtype or not _is_channels_last_dense(residual)):
        return False
    if running_mean is not None and (momentum is None or running_mean.dtype != torch.float32):
        return False
    return os.environ.get("FL4H_DETERMINISTIC", "0") != "1"


def batch_norm_act_reference(
    x: torch.Tensor, weight: torch.Tensor | None, bias: torch.Tensor | None, running_mean: torch.Tensor | None,
    running_var: torch.Tensor | None, training: bool, momentum: float, eps: float,
    residual: torch.Tensor | None = None, relu: bool = True,
) -> torch.Tensor:
    out = F_nn.batch_norm(x, running_mean, running_var, weight, bias, training, momentum, eps)
    if residual is not None:
        out = out + residual
    return F_nn.relu(out) if relu else out


class _BatchNormAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, weight, bias, running_mean, running_var, nbt, training, momentum, eps, relu,  # noqa: ANN001, ANN205
                presums=None, done=None):
        lib = _lib.load(True)
        n, c, h, w = x.shape
        m = n * h * w
        y = torch.empty_like(x)  # preserves channels-last strides
        stream = _lib.stream_ptr(x.device)
        is_bf16 = 1 if x.dtype == torch.bfloat16 else 0
        if training and presums is not None:
            # the producing convolution's epilogue already reduced sum / sum-of-squares (ops/conv.py): one streaming
            # apply pass, no statistics pass, no grid barrier; the kernel re-zeroes `presums` when it is done
            stats = torch.empty(4, c, dtype=torch.float32, device=x.device)
            err = lib.fl4h_bn_fwd_train_presum(
                _lib.ptr(x), _lib.ptr(residual), _lib.ptr(y), ctypes.c_int64(m), ctypes.c_int(c), _lib.ptr(presums),
                _lib.ptr(weight), _lib.ptr(bias), _lib.ptr(running_mean), _lib.ptr(running_var), _lib.ptr(nbt),
                ctypes.c_float(momentum if momentum is not None else 0.0), ctypes.c_float(eps), _lib.ptr(stats[0]),
                _lib.ptr(stats[1]), _lib.ptr(done), ctypes.c_int(is_bf16), ctypes.c_int(1 if relu else 0), stream,
            )
            _lib.check(err, "fl4h_bn_fwd_train_presum")
            _lib.count_launches(1)
            ctx.save_for_backward(x, y, weight, stats)
            ctx.relu, ctx.has_res, ctx.has_bias = relu, residual is not None, bias is not None
            ctx.training = True
            return y
        if training:
            acc, counter = _workspace(x.device, c)
            stats = torch.empty(4, c, dtype=torch.float32, device=x.device)  # mean, invstd, scale, shift
            err = lib.fl4h_bn_fwd_train(
                _lib.ptr(x), _lib.ptr(residual), _lib.ptr(y), ctypes.c_int64(m), ctypes.c_int(c), _lib.ptr(weight),
                _lib.ptr(bias), _lib.ptr(running_mean), _lib.ptr(running_var), _lib.ptr(nbt),
                ctypes.c_float(momentum if momentum is not None else 0.0), ctypes.c_float(eps), _lib.ptr(stats[0]),
                _lib.ptr(stats[1]), _lib.ptr(stats[2]), _lib.ptr(acc), _lib.ptr(counter), ctypes.c_int(is_bf16),
                ctypes.c_int(1 if relu else 0), ctypes.c_int(_allow_fused()), stream,
            )
            _lib.check(err, "fl4h_bn_fwd_train")
            _lib.count_launches(1 if _allow_fused() else 2)
            ctx.save_for_backward(x, y, weight, stats)
            ctx.relu, ctx.has_res, ctx.has_bias = relu, residual is not None, bias is not None
        else:
            err = lib.fl4h_bn_fwd_eval(
                _lib.ptr(x), _lib.ptr(residual), _lib.ptr(y), ctypes.c_int64(m), ctypes.c_int(c), _lib.ptr(weight),
                _lib.ptr(bias), _lib.ptr(running_mean), _lib.ptr(running_var), ctypes.c_float(eps), ctypes.c_int(is_bf16),
                ctypes.c_int(1 if relu else 0), stream,
            )
            _lib.check(err, "fl4h_bn_fwd_eval")
            _lib.count_launches(1)
            ctx.save_for_backward(x, y, weight, running_mean, running_var)
            ctx.relu, ctx.has_res, ctx.has_bias, ctx.eps = relu, residual is not None, bias is not None, eps
        ctx.training = training
        return y

    @staticmethod
    def backward(ctx, dy):  # noqa: ANN001, ANN205
        if not ctx.training:
            # eval-mode BN is an affine map: plain PyTorch math (rare path: frozen-statistics fine-tuning)
            x, y, weight, running_mean, running_var = ctx.saved_tensors
            g = dy * (y > 0).to(dy.dtype) if ctx.relu else dy
            invstd = torch.rsqrt(running_var + ctx.eps)
            scale = invstd * (weight if weight is not None else 1.0)
            dx = (g.float() * scale.view(1, -1, 1, 1)).to(x.dtype)
            xhat = (x.float() - running_mean.view(1, -1, 1, 1)) * invstd.view(1, -1, 1, 1)
            dweight = (g.float() * xhat).sum(dim=(0, 2, 3)) if weight is not None else None
            dbias = g.float().sum(dim=(0, 2, 3)) if ctx.has_bias else None
            return dx, (g if ctx.has_res else None), dweight, dbias, None, None, None, None, None, None, None, None, None
        x, y, weight, stats = ctx.saved_tensors
        lib = _lib.load(True)
        n, c, h, w = x.shape
        m = n * h * w
        if dy.dtype != x.dtype or not _is_channels_last_dense(dy):
            dy = dy.to(x.dtype).contiguous(memory_format=torch.channels_last)
        dx = torch.empty_like(x)
        dres = torch.empty_like(x) if ctx.has_res else None
        grads = torch.empty(2, c, dtype=torch.float32, device=x.device)
        coef = torch.empty(3, c, dtype=torch.float32, device=x.device)
        acc, counter = _workspace(x.device, c)
        err = lib.fl4h_bn_bwd(
            _lib.ptr(dy), _lib.ptr(y), _lib.ptr(x), ctypes.c_int64(m), ctypes.c_int(c), _lib.ptr(weight),
            _lib.ptr(stats[0]), _lib.ptr(stats[1]), _lib.ptr(dx), _lib.ptr(dres), _lib.ptr(grads[0]), _lib.ptr(grads[1]),
            _lib.ptr(coef), _lib.ptr(acc), _lib.ptr(counter), ctypes.c_int(1 if x.dtype == torch.bfloat16 else 0),
            ctypes.c_int(1 if ctx.relu else 0), ctypes.c_int(_allow_fused()), _lib.stream_ptr(x.device),
        )
        _lib.check(err, "fl4h_bn_bwd")
        _lib.count_launches(1 if _allow_fused() else 2)
        dweight = grads[0] if weight is not None else None
        dbias = grads[1] if ctx.has_bias else None
        return dx, dres, dweight, dbias, None, None, None, None, None, None, None, None, None


def batch_norm_act(
    x: torch.Tensor, weight: torch.Tensor | None, bias: torch.Tensor | None, running_mean: torch.Tensor | None,
    running_var: torch.Tensor | None, num_batches_tracked: torch.Tensor | None, training: bool, momentum: float | None,
    eps: float, residual: torch.Tensor | None = None, relu: bool = True,
    presums: torch.Tensor | None = None, done: torch.Tensor | None = None,
) -> torch.Tensor:
    """``relu(batch_norm(x) + residual)`` with BatchNorm2d semantics (running stats and the batch counter are updated
    in training mode).  ``presums`` ([2, C] fp32: per-channel sum / sum of squares of ``x``, e.g. from the convolution
    epilogue) + ``done`` (int32 election counter) skip the statistics pass; only valid when ``presums_eligible``."""
    use_batch_stats = training or running_mean is None
    if kernel_eligible(x, residual, momentum, use_batch_stats, running_mean):
        if weight is not None and weight.dtype != torch.float32:
            weight, bias = weight.float(), (bias.float() if bias is not None else None)
        if not use_batch_stats:
            presums = done = None
        return _BatchNormAct.apply(x, residual, weight, bias, running_mean, running_var,
                                   num_batches_tracked if use_batch_stats else None, use_batch_stats, momentum, eps, relu,
                                   presums, done)
    assert presums is None, "presums were produced but the BatchNorm kernel is not eligible: check presums_eligible first"
    if training and num_batches_tracked is not None:
        num_batches_tracked.add_(1)
    exp_factor = momentum
    if momentum is None:
        exp_factor = 1.0 / float(num_batches_tracked) if (training and num_batches_tracked is not None) else 0.0
    return batch_norm_act_reference(x, weight, bias, running_mean, running_var, use_batch_stats, exp_factor, eps, residual, relu)
