"""tcgen05 implicit-GEMM convolution (kernels: ``csrc/conv_tc.cu``): NHWC forward / data-gradient / weight-gradient.

Tensors are ordinary 4-D ``torch`` tensors in ``channels_last`` memory format (logical NCHW, physical NHWC); filters are
``[Cout, Cin, R, S]`` channels_last, i.e. physically ``[Cout, R, S, Cin]`` — exactly how 4-D parameters sit in the
``ParameterArena`` — so no layout conversion happens on the forward path.  fp32 tensors run TF32 tensor-core math (what
the reference's ``nn.Conv2d`` does under PyTorch's cuDNN defaults, ``examples/models/cnn_model.py:16-22``); bf16 tensors
run bf16.

The host side only describes the problem to the generic tap-GEMM kernel: which sub-lattices of the input exist
(strided layers address parity sub-lattices so every tap is a dense TMA box), which taps contribute, where their filter
columns start.  ``conv2d_reference`` is the stock composition (cuDNN / CPU) used as the numerics oracle and fallback.
"""

from __future__ import annotations

import ctypes
import os
from functools import lru_cache

import torch
import torch.nn.functional as F_nn

from fl4health_b200.ops import _lib

_ROW_BYTES = 128
_SM_COUNT = 148


def _is_cl(t: torch.Tensor) -> bool:
    return t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last)


def _pow2(n: int) -> bool:
    return n > 0 and (n & (n - 1)) == 0


def supported(x: torch.Tensor, weight: torch.Tensor, stride: int, padding: int, groups: int = 1, dilation: int = 1) -> bool:
    """Shapes the tcgen05 kernels cover: square 1x1 / 3x3 filters, stride 1 or 2, 'same'-style padding, power-of-two
    planes, channel counts that fill whole 128-byte TMA rows.  Anything else takes the stock path."""
    if not x.is_cuda or _lib.load() is None or groups != 1 or dilation != 1:
        return False
    if x.dtype not in (torch.float32, torch.bfloat16) or weight.dtype != x.dtype:
        return False
    if not _is_cl(x) or not _is_cl(weight):
        return False
    cout, cin, r, s = weight.shape
    n, _, h, w = x.shape
    if r != s or r not in (1, 3) or stride not in (1, 2) or padding != (r - 1) // 2:
        return False
    cpb = _ROW_BYTES // x.element_size()
    if cin % max(cpb, 64) != 0 or cout % 64 != 0:
        return False
    if h % stride or w % stride:
        return False
    ho, wo = h // stride, w // stride
    return _pow2(ho) and _pow2(wo) and wo <= 128 and ho * wo >= 16 and (ho * wo >= 64 or n % max(1, 64 // (ho * wo)) == 0)


def _arr(values: list[int], ctype=ctypes.c_int):  # noqa: ANN001, ANN202
    return (ctype * max(1, len(values)))(*values)


@lru_cache(maxsize=256)
def _forward_plan(h: int, w: int, cin: int, r: int, stride: int, pad: int) -> dict:
    """Input sub-lattices + tap table of a forward convolution (also used by the weight gradient)."""
    s = stride
    lattices = []  # (offset, hv, wv, sw, sh) in elements
    for ah in range(s):
        for aw in range(s):
            lattices.append(((ah * w + aw) * cin, (h - ah + s - 1) // s, (w - aw + s - 1) // s, s * cin, s * w * cin))
    dh, dw, mp, wcol = [], [], [], []
    for i in range(r):
        for j in range(r):
            qh, qw = i - pad, j - pad
            ah, aw = qh % s, qw % s
            dh.append((qh - ah) // s)
            dw.append((qw - aw) // s)
            mp.append(ah * s + aw)
            wcol.append((i * r + j) * cin)
    return {"lattices": lattices, "dh": dh, "dw": dw, "map": mp, "wcol": wcol}


def _pick_splits(ctas: int, k_blocks: int, limit: int) -> int:
    env = os.environ.get("FL4H_CONV_SPLITS")
    if env:
        return max(1, min(int(env), limit, k_blocks))
    if ctas >= 96:
        return 1
    want = max(1, _SM_COUNT // max(ctas, 1))
    for cand in (16, 8, 4, 2):
        if cand <= want and cand <= limit and k_blocks // cand >= 4:
            return cand
    return 1


def conv2d_forward(x: torch.Tensor, weight: torch.Tensor, stride: int, padding: int,
                   stats: torch.Tensor | None = None) -> torch.Tensor:
    """``y = conv2d(x, weight)`` (no bias).  ``stats`` ([2, Cout] fp32, zero on entry) receives per-channel sum and
    sum of squares of the stored outputs (BatchNorm's batch statistics, produced by the convolution epilogue)."""
    lib = _lib.load(True)
    n, cin, h, w = x.shape
    cout, _, r, _ = weight.shape
    ho, wo = h // stride, w // stride
    y = torch.empty((n, cout, ho, wo), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    plan = _forward_plan(h, w, cin, r, stride, padding)
    lat = plan["lattices"]
    cpb = _ROW_BYTES // x.element_size()
    m_tiles = (n * ho * wo + 127) // 128
    splits = _pick_splits(m_tiles * (cout // 64), len(plan["dh"]) * (cin // cpb), 8)
    err = lib.fl4h_conv_tap_gemm(
        _lib.ptr(x), _lib.ptr(weight), _lib.ptr(y), _lib.ptr(stats), ctypes.c_int(0 if x.dtype == torch.float32 else 1),
        ctypes.c_int(n), ctypes.c_int(cin), ctypes.c_int(cout), ctypes.c_int(r * r * cin), ctypes.c_int(len(lat)),
        _arr([l[0] for l in lat], ctypes.c_longlong), _arr([l[1] for l in lat]), _arr([l[2] for l in lat]),
        _arr([l[3] for l in lat], ctypes.c_longlong), _arr([l[4] for l in lat], ctypes.c_longlong),
        ctypes.c_longlong(h * w * cin), ctypes.c_int(ho), ctypes.c_int(wo), ctypes.c_longlong(0), ctypes.c_longlong(cout),
        ctypes.c_longlong(wo * cout), ctypes.c_longlong(ho * wo * cout), ctypes.c_int(len(plan["dh"])), _arr(plan["dh"]),
        _arr(plan["dw"]), _arr(plan["map"]), _arr(plan["wcol"]), ctypes.c_int(splits), _lib.stream_ptr(x.device),
    )
    _lib.check(err, "fl4h_conv_tap_gemm(forward)")
    _lib.count_launches(1)
    return y


def permute_filter_for_dgrad(weight: torch.Tensor) -> torch.Tensor:
    """[Cout, R, S, Cin] (physical) -> [Cin, R, S, Cout] (physical), returned as a channels_last [Cin, Cout, R, S] tensor:
    the K-major filter matrix of the data-gradient GEMM."""
    return weight.permute(1, 0, 2, 3).contiguous(memory_format=torch.channels_last)


def conv2d_dgrad(dy: torch.Tensor, weight_t: torch.Tensor, in_hw: tuple[int, int], stride: int, padding: int) -> torch.Tensor:
    """``dx`` of ``conv2d`` given ``weight_t = permute_filter_for_dgrad(weight)``."""
    lib = _lib.load(True)
    n, cout, ho, wo = dy.shape
    cin, _, r, _ = weight_t.shape
    h, w = in_hw
    dx = torch.empty((n, cin, h, w), dtype=dy.dtype, device=dy.device, memory_format=torch.channels_last)
    cpb = _ROW_BYTES // dy.element_size()
    dtype = ctypes.c_int(0 if dy.dtype == torch.float32 else 1)
    s = stride
    for ah in range(s):
        for aw in range(s):
            dh, dw, wcol = [], [], []
            for i in range(r):
                for j in range(r):
                    if (ah + padding - i) % s == 0 and (aw + padding - j) % s == 0:
                        dh.append((ah + padding - i) // s)
                        dw.append((aw + padding - j) // s)
                        wcol.append((i * r + j) * cout)
            hv, wv = h // s, w // s
            m_tiles = (n * hv * wv + 127) // 128
            splits = _pick_splits(m_tiles * (cin // 64), max(1, len(dh)) * (cout // cpb), 8) if dh else 1
            err = lib.fl4h_conv_tap_gemm(
                _lib.ptr(dy), _lib.ptr(weight_t), _lib.ptr(dx), None, dtype, ctypes.c_int(n), ctypes.c_int(cout), ctypes.c_int(cin),
                ctypes.c_int(r * r * cout), ctypes.c_int(1), _arr([0], ctypes.c_longlong), _arr([ho]), _arr([wo]),
                _arr([cout], ctypes.c_longlong), _arr([wo * cout], ctypes.c_longlong), ctypes.c_longlong(ho * wo * cout),
                ctypes.c_int(hv), ctypes.c_int(wv), ctypes.c_longlong((ah * w + aw) * cin), ctypes.c_longlong(s * cin),
                ctypes.c_longlong(s * w * cin), ctypes.c_longlong(h * w * cin), ctypes.c_int(len(dh)), _arr(dh), _arr(dw),
                _arr([0] * len(dh)), _arr(wcol), ctypes.c_int(splits), _lib.stream_ptr(dy.device),
            )
            _lib.check(err, "fl4h_conv_tap_gemm(dgrad)")
            _lib.count_launches(1)
    return dx


def conv2d_wgrad(x: torch.Tensor, dy: torch.Tensor, r: int, stride: int, padding: int) -> torch.Tensor:
    """``dW`` ([Cout, Cin, R, R] channels_last) of ``conv2d``: both operands read MN-major from the NHWC tensors."""
    lib = _lib.load(True)
    n, cin, h, w = x.shape
    _, cout, ho, wo = dy.shape
    dw_out = torch.empty((cout, cin, r, r), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    plan = _forward_plan(h, w, cin, r, stride, padding)
    lat = plan["lattices"]
    k_blocks = max(1, n * ho * wo // 64)
    ctas = ((cout + 127) // 128) * (cin // 64) * len(plan["dh"])
    splits = _pick_splits(ctas, k_blocks, 16)
    err = lib.fl4h_conv_wgrad(
        _lib.ptr(x), _lib.ptr(dy), _lib.ptr(dw_out), ctypes.c_int(0 if x.dtype == torch.float32 else 1), ctypes.c_int(n),
        ctypes.c_int(cin), ctypes.c_int(cout), ctypes.c_int(r * r * cin), ctypes.c_int(len(lat)),
        _arr([l[0] for l in lat], ctypes.c_longlong), _arr([l[1] for l in lat]), _arr([l[2] for l in lat]),
        _arr([l[3] for l in lat], ctypes.c_longlong), _arr([l[4] for l in lat], ctypes.c_longlong),
        ctypes.c_longlong(h * w * cin), ctypes.c_int(ho), ctypes.c_int(wo), ctypes.c_int(len(plan["dh"])), _arr(plan["dh"]),
        _arr(plan["dw"]), _arr(plan["map"]), _arr(plan["wcol"]), ctypes.c_int(splits), _lib.stream_ptr(x.device),
    )
    _lib.check(err, "fl4h_conv_wgrad")
    _lib.count_launches(1)
    return dw_out


def conv2d_reference(x: torch.Tensor, weight: torch.Tensor, stride: int, padding: int) -> torch.Tensor:
    return F_nn.conv2d(x, weight, None, stride, padding)


class _TcConv2d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, stride, padding, stats):  # noqa: ANN001, ANN205
        ctx.save_for_backward(x, weight)
        ctx.stride, ctx.padding = stride, padding
        return conv2d_forward(x, weight, stride, padding, stats)

    @staticmethod
    def backward(ctx, dy):  # noqa: ANN001, ANN205
        x, weight = ctx.saved_tensors
        if dy.dtype != x.dtype or not _is_cl(dy):
            dy = dy.to(x.dtype).contiguous(memory_format=torch.channels_last)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = conv2d_dgrad(dy, permute_filter_for_dgrad(weight), (x.shape[2], x.shape[3]), ctx.stride, ctx.padding)
        if ctx.needs_input_grad[1]:
            dw = conv2d_wgrad(x, dy, weight.shape[2], ctx.stride, ctx.padding)
        return dx, dw, None, None, None


def conv2d(x: torch.Tensor, weight: torch.Tensor, stride: int = 1, padding: int = 0,
           stats: torch.Tensor | None = None) -> torch.Tensor:
    """Differentiable tcgen05 convolution when the shape is covered, stock ``F.conv2d`` otherwise (``stats`` is then
    left untouched: callers check ``supported`` first when they rely on the epilogue statistics)."""
    if supported(x, weight, stride, padding):
        return _TcConv2d.apply(x, weight, stride, padding, stats)
    return conv2d_reference(x, weight, stride, padding)
