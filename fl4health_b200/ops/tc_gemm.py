"""Fused ``act(x @ W^T + b)`` on the tcgen05 tensor cores (kernel: ``csrc/tc_gemm.cu``).

``linear_bias_act(x, weight, bias, relu)`` runs the hand-written TMA -> tcgen05.mma -> TMEM pipeline for bf16 CUDA
operands (K-major activations and ``nn.Linear`` weights, which is how both are stored anyway) and falls back to
``torch.nn.functional.linear`` otherwise.  The autograd wrapper keeps the backward on library GEMMs: the forward --
the path validation and inference take, and the one that carries the bias/activation epilogue -- is the fused kernel.
"""

from __future__ import annotations

import ctypes
import os

import torch
import torch.nn.functional as F_nn

from fl4health_b200.ops import _lib


def kernel_eligible(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None) -> bool:
    if not (x.is_cuda and weight.is_cuda) or _lib.load() is None or os.environ.get("FL4H_TC_DISABLE") == "1":
        return False
    if x.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        return False
    if x.dim() < 2 or weight.dim() != 2 or not weight.is_contiguous():
        return False
    k, n = weight.shape[1], weight.shape[0]
    if x.shape[-1] != k or k % 8 != 0 or n % 8 != 0:  # 16-byte global strides for TMA / vector stores
        return False
    if x.data_ptr() % 16 or weight.data_ptr() % 16:
        return False
    return bias is None or (bias.is_cuda and bias.numel() == n)


ACTIVATIONS = {"none": 0, "relu": 1, "gelu": 2}


def _act_code(act: bool | str | None) -> int:
    if isinstance(act, bool) or act is None:
        return 1 if act else 0
    return ACTIVATIONS[act]


def linear_bias_act_reference(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None, relu: bool | str) -> torch.Tensor:
    out = F_nn.linear(x.float(), weight.float(), None if bias is None else bias.float())
    code = _act_code(relu)
    out = torch.relu(out) if code == 1 else (F_nn.gelu(out) if code == 2 else out)
    return out.to(x.dtype)


def pick_variant(m: int, n: int, k: int = 0) -> int:
    """0: one 128x128 tile per CTA; 1: persistent 128x128 (double-buffered TMEM); 2: persistent 128x256, 4 epilogue
    warps; 3: persistent 128x256, 8 epilogue warps; 4: CTA pair (``tcgen05.mma.cta_group::2``), 256x256 tile per pair,
    8 epilogue warps per CTA.  ``FL4H_TC_VARIANT`` forces one.  Measured on B200 with the calls captured in a CUDA graph
    (``benchmarks/tc_gemm_bench.py``): the pair kernel is the fastest or tied at every shape with N >= 256
    (903 vs 896 (v3) vs 798 (v2) TFLOP/s at 4096x2304x768; 1504 vs 1406 vs 1418 at 8192^3)."""
    forced = os.environ.get("FL4H_TC_VARIANT")
    if forced is not None:
        return int(forced)
    if n < 256:
        return 1
    return 3 if m <= 128 else 4


def _launch(x2d: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None, act: bool | int, variant: int | None = None,
            want_pre: bool = False) -> torch.Tensor | tuple[torch.Tensor, torch.Tensor]:
    lib = _lib.load(True)
    m, k = x2d.shape
    n = weight.shape[0]
    variant = pick_variant(m, n, k) if variant is None else variant
    out = torch.empty(m, n, dtype=torch.bfloat16, device=x2d.device)
    pre = torch.empty_like(out) if want_pre else None
    code = int(act)
    if bias is None or bias.dtype == torch.float32:
        bias_arg = bias
    elif bias.dtype == torch.bfloat16 and variant in (1, 2, 3, 4):  # master-weight mode: read the bf16 bias in the kernel
        bias_arg, code = bias, code | 0x100
    else:
        bias_arg = bias.float()
    if bias_arg is not None and not bias_arg.is_contiguous():
        bias_arg = bias_arg.contiguous()
    err = lib.fl4h_tc_linear_ex(
        _lib.ptr(x2d), _lib.ptr(weight), _lib.ptr(out), _lib.ptr(pre), _lib.ptr(bias_arg), ctypes.c_int(m), ctypes.c_int(n),
        ctypes.c_int(k), ctypes.c_int(code), ctypes.c_int(variant), _lib.stream_ptr(x2d.device),
    )
    if err != 0:
        raise RuntimeError(f"fl4h_tc_linear failed ({'CUresult ' + str(-err) if err < 0 else 'cudaError ' + str(err)})")
    _lib.count_launches(1)
    return (out, pre) if want_pre else out


class _LinearBiasAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, act):  # noqa: ANN001, ANN205
        x2d = x.reshape(-1, x.shape[-1])
        if not x2d.is_contiguous():
            x2d = x2d.contiguous()
        needs_grad = any(ctx.needs_input_grad[:3])
        if act == 2 and needs_grad:  # GELU backward needs the pre-activation: the epilogue stores it alongside
            out, pre = _launch(x2d, weight, bias, act, want_pre=True)
            ctx.save_for_backward(x2d, weight, pre)
        else:
            out = _launch(x2d, weight, bias, act)
            ctx.save_for_backward(x2d, weight, out if act == 1 else None)
        ctx.act, ctx.has_bias, ctx.in_shape = act, bias is not None, x.shape
        ctx.bias_dtype = None if bias is None else bias.dtype
        return out.reshape(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, grad_out):  # noqa: ANN001, ANN205
        x2d, weight, aux = ctx.saved_tensors
        g = grad_out.reshape(-1, grad_out.shape[-1])
        if ctx.act == 1:
            g = g * (aux > 0).to(g.dtype)
        elif ctx.act == 2:
            g = torch.ops.aten.gelu_backward(g.contiguous(), aux)
        grad_x = (g @ weight).reshape(ctx.in_shape) if ctx.needs_input_grad[0] else None
        grad_w = g.t() @ x2d if ctx.needs_input_grad[1] else None
        grad_b = None
        if ctx.has_bias and ctx.needs_input_grad[2]:  # one reduction kernel (fp32 accumulation inside) when dtypes agree
            grad_b = g.sum(dim=0) if ctx.bias_dtype == g.dtype else g.sum(dim=0, dtype=torch.float32).to(ctx.bias_dtype)
        return grad_x, grad_w, grad_b, None


def linear_bias_act(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None = None, relu: bool | str = False) -> torch.Tensor:
    """``relu``: False / True (ReLU) or one of ``"none" | "relu" | "gelu"`` (exact erf GELU)."""
    code = _act_code(relu)
    if kernel_eligible(x, weight, bias):
        return _LinearBiasAct.apply(x, weight, bias, code)
    out = F_nn.linear(x, weight, bias)
    return torch.relu(out) if code == 1 else (F_nn.gelu(out) if code == 2 else out)
