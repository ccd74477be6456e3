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
    """Split-K factor (cluster size along z): fill the machine at two resident CTAs per SM — with a 3-stage ring one CTA
    keeps ~72 KB in flight, i.e. ~80 GB/s at ~1 us TMA latency, so layers with few tiles are latency-bound until more
    CTAs share the reduction — while every split still has a few K blocks to pipeline."""
    env = os.environ.get("FL4H_CONV_SPLITS")
    if env:
        return max(1, min(int(env), limit, k_blocks))
    want = (2 * _SM_COUNT) // max(ctas, 1)
    for cand in (8, 4, 2):
        if cand <= want and cand <= limit and k_blocks // cand >= 3:
            return cand
    return 1


_MAX_TAPS = 9


def _tap_gemm(inp: torch.Tensor, wmat: torch.Tensor, out: torch.Tensor, stats: torch.Tensor | None, n: int, k_ch: int,
              out_ch: int, wmat_rows: int, wmat_cols: int, b_mn: bool, lattices: list, in_sn: int, plane: tuple[int, int],
              classes: list[dict], out_strides: tuple[int, int, int], splits: int, what: str) -> None:
    """Marshal one launch of ``fl4h_conv_tap_gemm``: ``classes`` = [{"off": out element offset, "dh": [...], "dw": [...],
    "map": [...], "wcol": [...]}, ...] (one per output sub-lattice)."""
    lib = _lib.load(True)

    def padded(key: str) -> list[int]:
        flat: list[int] = []
        for cls in classes:
            flat.extend(list(cls[key]) + [0] * (_MAX_TAPS - len(cls[key])))
        return flat

    err = lib.fl4h_conv_tap_gemm(
        _lib.ptr(inp), _lib.ptr(wmat), _lib.ptr(out), _lib.ptr(stats), ctypes.c_int(0 if inp.dtype == torch.float32 else 1),
        ctypes.c_int(n), ctypes.c_int(k_ch), ctypes.c_int(out_ch), ctypes.c_int(wmat_rows), ctypes.c_int(wmat_cols),
        ctypes.c_int(1 if b_mn else 0), ctypes.c_int(len(lattices)), _arr([l[0] for l in lattices], ctypes.c_longlong),
        _arr([l[1] for l in lattices]), _arr([l[2] for l in lattices]), _arr([l[3] for l in lattices], ctypes.c_longlong),
        _arr([l[4] for l in lattices], ctypes.c_longlong), ctypes.c_longlong(in_sn), ctypes.c_int(plane[0]), ctypes.c_int(plane[1]),
        ctypes.c_int(len(classes)), _arr([cls["off"] for cls in classes], ctypes.c_longlong), ctypes.c_longlong(out_strides[0]),
        ctypes.c_longlong(out_strides[1]), ctypes.c_longlong(out_strides[2]), _arr([len(cls["dh"]) for cls in classes]),
        _arr(padded("dh")), _arr(padded("dw")), _arr(padded("map")), _arr(padded("wcol")), ctypes.c_int(splits),
        _lib.stream_ptr(inp.device),
    )
    _lib.check(err, f"fl4h_conv_tap_gemm({what})")
    _lib.count_launches(1)


def conv2d_forward(x: torch.Tensor, weight: torch.Tensor, stride: int, padding: int,
                   stats: torch.Tensor | None = None) -> torch.Tensor:
    """``y = conv2d(x, weight)`` (no bias).  ``stats`` ([2, Cout] fp32, zero on entry) receives per-channel sum and
    sum of squares of the stored outputs (BatchNorm's batch statistics, produced by the convolution epilogue)."""
    n, cin, h, w = x.shape
    cout, _, r, _ = weight.shape
    ho, wo = h // stride, w // stride
    y = torch.empty((n, cout, ho, wo), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    plan = _forward_plan(h, w, cin, r, stride, padding)
    cpb = _ROW_BYTES // x.element_size()
    m_tiles = (n * ho * wo + 127) // 128
    splits = _pick_splits(m_tiles * (cout // 64), len(plan["dh"]) * (cin // cpb), 8)
    classes = [{"off": 0, "dh": plan["dh"], "dw": plan["dw"], "map": plan["map"], "wcol": plan["wcol"]}]
    _tap_gemm(x, weight, y, stats, n, cin, cout, cout, r * r * cin, False, plan["lattices"], h * w * cin, (ho, wo), classes,
              (cout, wo * cout, ho * wo * cout), splits, "forward")
    return y


@lru_cache(maxsize=256)
def _dgrad_classes(h: int, w: int, cin: int, r: int, stride: int, pad: int) -> tuple:
    """Output sub-lattices (one per parity of a strided layer) and their taps for the data gradient
    ``dx[n, h, w] = sum_{r,s} dy[n, (h + pad - r) / stride, (w + pad - s) / stride] . W[:, r, s, :]`` (exact divisions only)."""
    s = stride
    classes = []
    for ah in range(s):
        for aw in range(s):
            dh, dw, wcol = [], [], []
            for i in range(r):
                for j in range(r):
                    if (ah + pad - i) % s == 0 and (aw + pad - j) % s == 0:
                        dh.append((ah + pad - i) // s)
                        dw.append((aw + pad - j) // s)
                        wcol.append((i * r + j) * cin)  # column of tap (i, j)'s input-channel block in W[Cout, R*S*Cin]
            classes.append({"off": (ah * w + aw) * cin, "dh": dh, "dw": dw, "map": [0] * len(dh), "wcol": wcol})
    return tuple(classes)


def conv2d_dgrad(dy: torch.Tensor, weight: torch.Tensor, in_hw: tuple[int, int], stride: int, padding: int) -> torch.Tensor:
    """``dx`` of ``conv2d``: ONE launch (the output parities of a strided layer are grid classes); the filters are read
    MN-major from the forward layout ``[Cout, R, S, Cin]`` — no flipped / transposed copy is made."""
    n, cout, ho, wo = dy.shape
    _, cin, r, _ = weight.shape
    h, w = in_hw
    dx = torch.empty((n, cin, h, w), dtype=dy.dtype, device=dy.device, memory_format=torch.channels_last)
    cpb = _ROW_BYTES // dy.element_size()
    s = stride
    classes = list(_dgrad_classes(h, w, cin, r, stride, padding))
    hv, wv = h // s, w // s
    m_tiles = (n * hv * wv + 127) // 128
    fewest = min((len(c["dh"]) for c in classes if c["dh"]), default=1)
    splits = _pick_splits(m_tiles * (cin // 64) * len(classes), fewest * (cout // cpb), 8)
    _tap_gemm(dy, weight, dx, None, n, cout, cin, cout, r * r * cin, True, [(0, ho, wo, cout, wo * cout)], ho * wo * cout,
              (hv, wv), classes, (s * cin, s * w * cin, h * w * cin), splits, "dgrad")
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
    splits = _pick_splits(ctas, k_blocks, 8)
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


# ------------------------------------------------------------------------------------------------------------------
# stem (Cin <= 4): CUDA-core kernels, csrc/conv_stem.cu
# ------------------------------------------------------------------------------------------------------------------
_STEM_SCRATCH: dict[tuple, tuple[torch.Tensor, torch.Tensor]] = {}


def stem_supported(x: torch.Tensor, weight: torch.Tensor, stride: int, padding: int, groups: int = 1, dilation: int = 1) -> bool:
    """3x3 / stride 1 / pad 1 / 64 filters over 1, 3 or 4 input channels (the first layer of the image models)."""
    if not x.is_cuda or _lib.load() is None or groups != 1 or dilation != 1 or stride != 1 or padding != 1:
        return False
    if x.dtype not in (torch.float32, torch.bfloat16) or weight.dtype != x.dtype or not _is_cl(x) or not _is_cl(weight):
        return False
    cout, cin, r, s = weight.shape
    return (cout == 64 and cin in (1, 3, 4) and r == 3 and s == 3 and x.shape[2] % 8 == 0 and x.shape[3] % 32 == 0
            and x.shape[3] <= 64)


def stem_forward(x: torch.Tensor, weight: torch.Tensor, stats: torch.Tensor | None = None) -> torch.Tensor:
    lib = _lib.load(True)
    n, cin, h, w = x.shape
    y = torch.empty((n, 64, h, w), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    err = lib.fl4h_conv_stem_fwd(_lib.ptr(x), _lib.ptr(weight), _lib.ptr(y), _lib.ptr(stats),
                                 ctypes.c_int(0 if x.dtype == torch.float32 else 1), ctypes.c_int(n), ctypes.c_int(h),
                                 ctypes.c_int(w), ctypes.c_int(cin), ctypes.c_int(64), _lib.stream_ptr(x.device))
    _lib.check(err, "fl4h_conv_stem_fwd")
    _lib.count_launches(1)
    return y


def stem_wgrad(x: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    lib = _lib.load(True)
    n, cin, h, w = x.shape
    dw_out = torch.empty((64, cin, 3, 3), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    key = (x.device.index, torch.cuda.current_stream(x.device).cuda_stream, cin)
    scratch = _STEM_SCRATCH.get(key)
    if scratch is None:  # self-resetting accumulators: one pair per (device, stream, Cin)
        scratch = (torch.zeros(64 * 9 * cin, dtype=torch.float32, device=x.device), torch.zeros(1, dtype=torch.int32, device=x.device))
        _STEM_SCRATCH[key] = scratch
    err = lib.fl4h_conv_stem_wgrad(_lib.ptr(x), _lib.ptr(dy), _lib.ptr(dw_out), _lib.ptr(scratch[0]), _lib.ptr(scratch[1]),
                                   ctypes.c_int(0 if x.dtype == torch.float32 else 1), ctypes.c_int(n), ctypes.c_int(h),
                                   ctypes.c_int(w), ctypes.c_int(cin), ctypes.c_int(64), _lib.stream_ptr(x.device))
    _lib.check(err, "fl4h_conv_stem_wgrad")
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
            dx = conv2d_dgrad(dy, weight, (x.shape[2], x.shape[3]), ctx.stride, ctx.padding)
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
