"""Drop-in layers backed by the fused sm_100a kernels (same parameters / buffers / state-dict keys as the stock
``torch.nn`` layers they subclass, so checkpoints and parameter exchange are unaffected)."""

from __future__ import annotations

import os

import torch
from torch import nn

from fl4health_b200.engine import streams
from fl4health_b200.ops import conv as tc_conv
from fl4health_b200.ops.layer_norm import add_dropout_layer_norm
from fl4health_b200.ops.bn_act import batch_norm_act, presums_eligible
from fl4health_b200.ops.tc_gemm import linear_bias_act


class ResidualLayerNorm(nn.LayerNorm):
    """``LayerNorm(residual + dropout(y))`` -- a transformer sub-layer's epilogue -- as one kernel per direction on CUDA
    (``ops/layer_norm.py``); exactly ``nn.LayerNorm`` (same parameters / state-dict) applied to the stock composition
    everywhere else.  ``dropout`` is the probability applied to ``y`` in training mode."""

    def __init__(self, normalized_shape: int, eps: float = 1e-5, dropout: float = 0.0) -> None:
        super().__init__(normalized_shape, eps=eps)
        self.dropout = dropout

    def forward(self, y: torch.Tensor, residual: torch.Tensor | None = None) -> torch.Tensor:  # type: ignore[override]
        return add_dropout_layer_norm(y, residual, self.weight, self.bias, self.eps, self.dropout, self.training)

    def extra_repr(self) -> str:
        return super().extra_repr() + f", dropout={self.dropout}"


class BatchNormAct2d(nn.BatchNorm2d):
    """``relu(bn(x) + residual)`` in two kernel launches forward / two backward on channels-last CUDA tensors
    (``ops/csrc/bn_act.cu``); exactly ``nn.BatchNorm2d`` (+ add + relu) everywhere else."""

    def __init__(self, num_features: int, eps: float = 1e-5, momentum: float | None = 0.1, affine: bool = True,
                 track_running_stats: bool = True, relu: bool = True, device=None, dtype=None) -> None:  # noqa: ANN001
        super().__init__(num_features, eps, momentum, affine, track_running_stats, device=device, dtype=dtype)
        self.relu = relu

    def forward(self, x: torch.Tensor, residual: torch.Tensor | None = None,  # type: ignore[override]
                presums: torch.Tensor | None = None) -> torch.Tensor:
        self._check_input_dim(x)
        return batch_norm_act(
            x, self.weight, self.bias, self.running_mean if self.track_running_stats else None,
            self.running_var if self.track_running_stats else None,
            self.num_batches_tracked if self.track_running_stats else None, self.training, self.momentum, self.eps,
            residual=residual, relu=self.relu, presums=presums, done=self._presum_done if presums is not None else None,
        )

    # statistics produced by the preceding convolution's epilogue (ops/conv.py): a self-resetting [2, C] accumulator
    # and the election counter of the apply kernel; plain attributes, deliberately NOT buffers (not model state)
    _presum_buf: torch.Tensor | None = None
    _presum_done: torch.Tensor | None = None

    def presum_buffer(self, like: torch.Tensor, residual: torch.Tensor | None) -> torch.Tensor | None:
        """The accumulator the producing convolution should reduce into, or None when this layer will not take the
        statistics-free kernel path for this input (eval mode, unsupported dtype, deterministic mode ...)."""
        use_batch_stats = self.training or not self.track_running_stats
        if not presums_eligible(like, self.num_features, residual, self.momentum, use_batch_stats,
                                self.running_mean if self.track_running_stats else None):
            return None
        if self._presum_buf is None or self._presum_buf.device != like.device:
            self._presum_buf = torch.zeros(2, self.num_features, dtype=torch.float32, device=like.device)
            self._presum_done = torch.zeros(1, dtype=torch.int32, device=like.device)
        return self._presum_buf

    def extra_repr(self) -> str:
        return super().extra_repr() + f", relu={self.relu}"


def bn_act(bn: nn.Module, x: torch.Tensor, residual: torch.Tensor | None = None, relu: bool = True) -> torch.Tensor:
    """Apply ``bn`` then (+residual)(relu).  Works for any normalisation module (e.g. the GroupNorm a DP validator
    swapped in); the fused path is taken when ``bn`` is a ``BatchNormAct2d``."""
    if isinstance(bn, BatchNormAct2d):
        assert bn.relu == relu
        return bn(x, residual)
    out = bn(x)
    if residual is not None:
        out = out + residual
    return torch.relu(out) if relu else out


def conv_bn_act(conv: nn.Module, bn: nn.Module, x: torch.Tensor, residual: torch.Tensor | None = None,
                relu: bool = True) -> torch.Tensor:
    """``relu(bn(conv(x)) + residual)``.  When ``conv`` runs on the tcgen05 kernel and ``bn`` is a ``BatchNormAct2d`` in
    training mode, the convolution epilogue reduces the batch statistics and BatchNorm becomes a single apply pass."""
    if isinstance(conv, TcConv2d) and isinstance(bn, BatchNormAct2d):
        if torch.is_autocast_enabled() and x.is_cuda and x.dtype != conv.weight.dtype == torch.bfloat16:
            x = x.to(torch.bfloat16)
    if isinstance(conv, TcConv2d) and isinstance(bn, BatchNormAct2d) and conv.kernel_applies(x):
        sums = bn.presum_buffer(x, residual)
        if sums is not None:
            assert bn.relu == relu
            return bn(conv(x, stats=sums), residual, presums=sums)
    return bn_act(bn, conv(x), residual, relu)


def _wgrad_may_overlap(weight: torch.Tensor) -> bool:
    """The weight gradient may be produced on the side stream only when autograd will ASSIGN it (``.grad`` is None:
    table-gradient / bf16-shadow engines).  With a live ``.grad`` (flat gradient region, gradient accumulation)
    AccumulateGrad adds it in place on the main stream right after this node returns, before the deferred join."""
    return streams.overlap_enabled() and weight.grad is None


class _TcConvFn(torch.autograd.Function):
    """tcgen05 implicit-GEMM convolution (``ops/conv.py``); the weight gradient runs on a side stream like the stock
    path's (only the optimizer needs it), the data gradient stays on the critical path."""

    @staticmethod
    def forward(ctx, x, weight, stride, padding, stats):  # noqa: ANN001, ANN205
        ctx.save_for_backward(x, weight)
        ctx.conf = (stride, padding)
        return tc_conv.conv2d_forward(x, weight, stride, padding, stats)

    @staticmethod
    def backward(ctx, grad_out):  # noqa: ANN001, ANN205
        x, weight = ctx.saved_tensors
        stride, padding = ctx.conf
        if grad_out.dtype != x.dtype or not grad_out.is_contiguous(memory_format=torch.channels_last):
            grad_out = grad_out.to(x.dtype).contiguous(memory_format=torch.channels_last)
        grad_x = grad_w = None
        if ctx.needs_input_grad[1]:
            if _wgrad_may_overlap(weight):
                main = torch.cuda.current_stream(x.device)
                side = streams.fork(x.device)
                with torch.cuda.stream(side):
                    grad_w = tc_conv.conv2d_wgrad(x, grad_out, weight.shape[2], stride, padding)
                grad_w.record_stream(main)
                streams.defer_join(x.device, grad_out, x, weight)
            else:
                grad_w = tc_conv.conv2d_wgrad(x, grad_out, weight.shape[2], stride, padding)
        if ctx.needs_input_grad[0]:
            grad_x = tc_conv.conv2d_dgrad(grad_out, weight, (x.shape[2], x.shape[3]),
                                          stride, padding)
        return grad_x, grad_w, None, None, None


class _StemConvFn(torch.autograd.Function):
    """First-layer convolution over raw image channels (``ops/csrc/conv_stem.cu``); the input never needs a gradient."""

    @staticmethod
    def forward(ctx, x, weight, stats):  # noqa: ANN001, ANN205
        ctx.save_for_backward(x, weight)
        return tc_conv.stem_forward(x, weight, stats)

    @staticmethod
    def backward(ctx, grad_out):  # noqa: ANN001, ANN205
        x, weight = ctx.saved_tensors
        if ctx.needs_input_grad[0]:
            raise RuntimeError("the stem convolution kernel does not produce input gradients (FL4H_TC_CONV=0 for that)")
        if grad_out.dtype != x.dtype or not grad_out.is_contiguous(memory_format=torch.channels_last):
            grad_out = grad_out.to(x.dtype).contiguous(memory_format=torch.channels_last)
        if not _wgrad_may_overlap(weight):
            return None, tc_conv.stem_wgrad(x, grad_out), None
        main = torch.cuda.current_stream(x.device)
        side = streams.fork(x.device)
        with torch.cuda.stream(side):
            grad_w = tc_conv.stem_wgrad(x, grad_out)
        grad_w.record_stream(main)
        streams.defer_join(x.device, grad_out, x)
        return None, grad_w, None


class _ConvOverlappedWgrad(torch.autograd.Function):
    """``conv2d`` whose backward issues the data gradient on the current stream (critical path) and the weight
    gradient on a side stream (only the optimizer needs it); see ``engine/streams.py``."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding, dilation, groups):  # noqa: ANN001, ANN205
        ctx.save_for_backward(x, weight)
        ctx.conf = (stride, padding, dilation, groups, None if bias is None else list(bias.shape))
        return torch.nn.functional.conv2d(x, weight, bias, stride, padding, dilation, groups)

    @staticmethod
    def backward(ctx, grad_out):  # noqa: ANN001, ANN205
        x, weight = ctx.saved_tensors
        stride, padding, dilation, groups, bias_sizes = ctx.conf
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], bias_sizes is not None and ctx.needs_input_grad[2]
        zeros = [0] * len(stride)
        grad_w = grad_b = grad_x = None
        if (need_w or need_b) and not _wgrad_may_overlap(weight):
            _, grad_w, grad_b = torch.ops.aten.convolution_backward(
                grad_out, x, weight, bias_sizes, stride, padding, dilation, False, zeros, groups, [False, need_w, need_b])
        elif need_w or need_b:
            main = torch.cuda.current_stream(x.device)
            side = streams.fork(x.device)
            with torch.cuda.stream(side):
                _, grad_w, grad_b = torch.ops.aten.convolution_backward(
                    grad_out, x, weight, bias_sizes, stride, padding, dilation, False, zeros, groups, [False, need_w, need_b])
            for g in (grad_w, grad_b):
                if g is not None:
                    g.record_stream(main)  # produced on the side stream, consumed (and freed) on the main stream
            streams.defer_join(x.device, grad_out, x, weight)
        if need_x:
            grad_x, _, _ = torch.ops.aten.convolution_backward(
                grad_out, x, weight, bias_sizes, stride, padding, dilation, False, zeros, groups, [True, False, False])
        return grad_x, grad_w, grad_b, None, None, None, None


class Conv2dOverlapWgrad(nn.Conv2d):
    """``nn.Conv2d`` (same parameters / state-dict) whose weight-gradient kernel overlaps the rest of the backward
    pass on CUDA.  Falls back to the stock op whenever the fast path's assumptions do not hold."""

    def forward(self, input: torch.Tensor) -> torch.Tensor:  # noqa: A002
        if (
            input.is_cuda and torch.is_grad_enabled() and self.padding_mode == "zeros" and streams.overlap_enabled()
            and not isinstance(self.padding, str) and input.dtype == self.weight.dtype
            and (self.weight.requires_grad or input.requires_grad)
        ):
            return _ConvOverlappedWgrad.apply(input, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        return super().forward(input)


class TcConv2d(Conv2dOverlapWgrad):
    """``nn.Conv2d`` (same parameters / state-dict) whose forward, data gradient and weight gradient are the hand-written
    tcgen05 implicit-GEMM kernels (``ops/csrc/conv_tc.cu``) for the shapes they cover — bias-free square 1x1 / 3x3 filters,
    stride 1 / 2, channels_last fp32 (TF32 math) or bf16 — and ``Conv2dOverlapWgrad`` (cuDNN) for everything else.
    ``FL4H_TC_CONV=0`` forces the library path (A/B runs)."""

    def _plain(self, x: torch.Tensor) -> bool:
        if os.environ.get("FL4H_TC_CONV", "1") == "0" or self.bias is not None or self.padding_mode != "zeros":
            return False
        if isinstance(self.padding, str) or self.stride[0] != self.stride[1] or self.padding[0] != self.padding[1]:
            return False
        return x.dtype == self.weight.dtype

    def stem_applies(self, x: torch.Tensor) -> bool:
        """Raw-image first layer (Cin <= 4): the CUDA-core stem kernels; the input must not require a gradient."""
        return (self._plain(x) and not x.requires_grad and os.environ.get("FL4H_STEM", "1") != "0"
                and tc_conv.stem_supported(x, self.weight, self.stride[0], self.padding[0], self.groups, self.dilation[0]))

    def kernel_applies(self, x: torch.Tensor) -> bool:
        if not self._plain(x):
            return False
        return (tc_conv.supported(x, self.weight, self.stride[0], self.padding[0], self.groups, self.dilation[0])
                or self.stem_applies(x))

    def forward(self, input: torch.Tensor, stats: torch.Tensor | None = None) -> torch.Tensor:  # noqa: A002
        if torch.is_autocast_enabled() and input.is_cuda and input.dtype != self.weight.dtype == torch.bfloat16:
            input = input.to(torch.bfloat16)  # master-weight mode: weights already bf16, the image batch follows autocast
        if self.stem_applies(input):
            return _StemConvFn.apply(input, self.weight, stats)
        if self.kernel_applies(input):
            return _TcConvFn.apply(input, self.weight, self.stride[0], self.padding[0], stats)
        assert stats is None, "epilogue statistics requested for a shape the tcgen05 kernel does not cover"
        return super().forward(input)


class LinearAct(nn.Linear):
    """``nn.Linear`` (+ optional fused ReLU / GELU) whose bf16 CUDA forward is the hand-written tcgen05 / TMEM / TMA GEMM
    with the bias + activation epilogue (``ops/csrc/tc_gemm.cu``); identical parameters and state-dict keys.
    ``activation``: ``None | "relu" | "gelu"`` (``relu=True`` is the older spelling of ``activation="relu"``)."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True, relu: bool = False, device=None, dtype=None,  # noqa: ANN001
                 activation: str | None = None) -> None:
        super().__init__(in_features, out_features, bias, device=device, dtype=dtype)
        assert activation in (None, "none", "relu", "gelu")
        self.relu = "relu" if relu else (activation or "none")

    def forward(self, input: torch.Tensor) -> torch.Tensor:  # noqa: A002
        if torch.is_autocast_enabled() and input.is_cuda and input.dtype != self.weight.dtype == torch.bfloat16:
            input = input.to(torch.bfloat16)  # master-weight mode: weights already bf16, activations follow autocast
        if self._use_kernel(input):
            return linear_bias_act(input, self.weight, self.bias, self.relu)
        out = nn.functional.linear(input, self.weight, self.bias)
        return torch.relu(out) if self.relu == "relu" else (nn.functional.gelu(out) if self.relu == "gelu" else out)

    def _use_kernel(self, input: torch.Tensor) -> bool:  # noqa: A002
        """``FL4H_TC_LINEAR``: ``always`` | ``never`` | ``auto`` (default).  ``auto`` uses the hand-written kernel where
        it is the faster choice on B200 (``profiles/README.md``, tcgen05 table, CUDA-graph timed):

        * fused **ReLU** on any layer with >= 256 outputs and >= 1024 rows: 903 vs 681 TFLOP/s at 4096x2304x768,
          1341 vs 839 at 16384x4096x1024, 1400 vs 1349 at 4096^3 against library GEMM + an elementwise pass;
        * fused **GELU** from 4096^3 of work up (1294 vs 1119): below that the erf chain makes the epilogue longer than
          a K = 768 mainloop and library GEMM + GELU pass wins (623 vs 493 at 4096x3072x768);
        * plain linears stay on the library (this kernel reaches 84-100 % of it: 903 vs 1073 at 4096x2304x768, 1504 vs
          1490 at 8192^3)."""
        policy = os.environ.get("FL4H_TC_LINEAR", "auto")
        if policy == "never" or not input.is_cuda:
            return False
        if policy == "always":
            return True
        rows = input.numel() // max(input.shape[-1], 1)
        if self.relu == "relu":
            return self.out_features >= 256 and rows >= 1024
        return self.relu == "gelu" and rows * self.in_features * self.out_features >= 4096 ** 3

    def extra_repr(self) -> str:
        return super().extra_repr() + f", activation={self.relu}"
