"""FedPM masked layers: frozen weights + learnable scores; forward = stock op on ``Bernoulli(sigmoid(score)) * weight``.

Parity: ``fl4health/model_bases/masked_layers/{masked_linear.py:11-104, masked_conv.py:15-878,
masked_normalization_layers.py:19-321}`` — same class names, ``weight_scores`` / ``bias_scores`` parameter names and
``from_pretrained`` constructors.  One generic implementation drives every layer type (the reference spells each of the
11 classes out by hand); mask sampling is the fused ``ops.masked`` kernel on CUDA.
"""

from __future__ import annotations

from typing import Any

import torch
import torch.nn.functional as F
from torch import Tensor, nn
from torch.nn.parameter import Parameter

from fl4health_b200.ops.masked import masked_parameter


class _MaskedParamsMixin:
    """Adds score parameters next to frozen ``weight`` / ``bias`` and samples masked versions of them."""

    weight: Any
    bias: Any

    def _init_scores(self) -> None:
        if self.weight is not None:
            self.weight.requires_grad = False
            self.weight_scores = Parameter(torch.randn_like(self.weight), requires_grad=True)
        else:
            self.register_parameter("weight_scores", None)  # type: ignore[attr-defined]
        if self.bias is not None:
            self.bias.requires_grad = False
            self.bias_scores = Parameter(torch.randn_like(self.bias), requires_grad=True)
        else:
            self.register_parameter("bias_scores", None)  # type: ignore[attr-defined]

    def masked_weight(self) -> Tensor | None:
        return masked_parameter(self.weight_scores, self.weight) if self.weight is not None else None

    def masked_bias(self) -> Tensor | None:
        return masked_parameter(self.bias_scores, self.bias) if self.bias is not None else None

    def _adopt(self, source: nn.Module) -> None:
        """Copy the pretrained (frozen) parameters of ``source`` and draw fresh scores."""
        if getattr(source, "weight", None) is not None:
            self.weight = Parameter(source.weight.clone().detach(), requires_grad=False)
            self.weight_scores = Parameter(torch.randn_like(source.weight), requires_grad=True)
        if getattr(source, "bias", None) is not None:
            self.bias = Parameter(source.bias.clone().detach(), requires_grad=False)
            self.bias_scores = Parameter(torch.randn_like(source.bias), requires_grad=True)


class MaskedLinear(_MaskedParamsMixin, nn.Linear):
    def __init__(self, in_features: int, out_features: int, bias: bool = True, device: Any = None, dtype: Any = None) -> None:
        super().__init__(in_features, out_features, bias, device, dtype)
        self._init_scores()

    def forward(self, input: Tensor) -> Tensor:
        return F.linear(input, self.masked_weight(), self.masked_bias())

    @classmethod
    def from_pretrained(cls, linear_module: nn.Linear) -> MaskedLinear:
        module = cls(linear_module.in_features, linear_module.out_features, bias=linear_module.bias is not None)
        module._adopt(linear_module)
        return module


_CONV_KEYS = ("in_channels", "out_channels", "kernel_size", "stride", "padding", "dilation", "groups", "padding_mode")


def _make_masked_conv(base: type, name: str) -> type:
    class _MaskedConv(_MaskedParamsMixin, base):  # type: ignore[misc, valid-type]
        def __init__(self, *args: Any, **kwargs: Any) -> None:
            super().__init__(*args, **kwargs)
            self._init_scores()

        def forward(self, input: Tensor) -> Tensor:
            return self._conv_forward(input, self.masked_weight(), self.masked_bias())

        @classmethod
        def from_pretrained(cls, conv_module: nn.Module) -> Any:
            kwargs = {key: getattr(conv_module, key) for key in _CONV_KEYS}
            module = cls(bias=conv_module.bias is not None, **kwargs)
            module._adopt(conv_module)
            return module

    _MaskedConv.__name__ = _MaskedConv.__qualname__ = name
    return _MaskedConv


def _make_masked_conv_transpose(base: type, name: str, functional: Any, num_spatial_dims: int) -> type:
    class _MaskedConvTranspose(_MaskedParamsMixin, base):  # type: ignore[misc, valid-type]
        def __init__(self, *args: Any, **kwargs: Any) -> None:
            super().__init__(*args, **kwargs)
            self._init_scores()

        def forward(self, input: Tensor, output_size: list[int] | None = None) -> Tensor:
            if self.padding_mode != "zeros":
                raise ValueError(f"Only `zeros` padding mode is supported for {name}")
            output_padding = self._output_padding(
                input, output_size, self.stride, self.padding, self.kernel_size, num_spatial_dims, self.dilation
            )
            return functional(input, self.masked_weight(), self.masked_bias(), self.stride, self.padding,
                              output_padding, self.groups, self.dilation)

        @classmethod
        def from_pretrained(cls, conv_module: nn.Module) -> Any:
            kwargs = {key: getattr(conv_module, key) for key in _CONV_KEYS}
            module = cls(bias=conv_module.bias is not None, output_padding=conv_module.output_padding, **kwargs)
            module._adopt(conv_module)
            return module

    _MaskedConvTranspose.__name__ = _MaskedConvTranspose.__qualname__ = name
    return _MaskedConvTranspose


MaskedConv1d = _make_masked_conv(nn.Conv1d, "MaskedConv1d")
MaskedConv2d = _make_masked_conv(nn.Conv2d, "MaskedConv2d")
MaskedConv3d = _make_masked_conv(nn.Conv3d, "MaskedConv3d")
MaskedConvTranspose1d = _make_masked_conv_transpose(nn.ConvTranspose1d, "MaskedConvTranspose1d", F.conv_transpose1d, 1)
MaskedConvTranspose2d = _make_masked_conv_transpose(nn.ConvTranspose2d, "MaskedConvTranspose2d", F.conv_transpose2d, 2)
MaskedConvTranspose3d = _make_masked_conv_transpose(nn.ConvTranspose3d, "MaskedConvTranspose3d", F.conv_transpose3d, 3)


class MaskedLayerNorm(_MaskedParamsMixin, nn.LayerNorm):
    def __init__(self, normalized_shape: Any, eps: float = 1e-5, elementwise_affine: bool = True, bias: bool = True,
                 device: Any = None, dtype: Any = None) -> None:
        super().__init__(normalized_shape, eps=eps, elementwise_affine=elementwise_affine, bias=bias, device=device, dtype=dtype)
        self._init_scores()

    def forward(self, input: Tensor) -> Tensor:
        return F.layer_norm(input, self.normalized_shape, self.masked_weight(), self.masked_bias(), self.eps)

    @classmethod
    def from_pretrained(cls, layer_norm_module: nn.LayerNorm) -> MaskedLayerNorm:
        module = cls(layer_norm_module.normalized_shape, eps=layer_norm_module.eps,
                     elementwise_affine=layer_norm_module.elementwise_affine, bias=layer_norm_module.bias is not None)
        module._adopt(layer_norm_module)
        return module


class _MaskedBatchNorm(_MaskedParamsMixin, nn.modules.batchnorm._BatchNorm):
    def __init__(self, num_features: int, eps: float = 1e-5, momentum: float | None = 0.1, affine: bool = True,
                 track_running_stats: bool = True, device: Any = None, dtype: Any = None) -> None:
        super().__init__(num_features, eps, momentum, affine, track_running_stats, device=device, dtype=dtype)
        self._init_scores()

    def forward(self, input: Tensor) -> Tensor:
        self._check_input_dim(input)
        factor = 0.0 if self.momentum is None else self.momentum
        if self.training and self.track_running_stats and self.num_batches_tracked is not None:
            self.num_batches_tracked.add_(1)
            factor = 1.0 / float(self.num_batches_tracked) if self.momentum is None else self.momentum
        use_batch_stats = self.training or (self.running_mean is None and self.running_var is None)
        return F.batch_norm(
            input,
            self.running_mean if not self.training or self.track_running_stats else None,
            self.running_var if not self.training or self.track_running_stats else None,
            self.masked_weight(), self.masked_bias(), use_batch_stats, factor, self.eps,
        )

    @classmethod
    def from_pretrained(cls, batch_norm_module: nn.modules.batchnorm._BatchNorm) -> Any:
        module = cls(batch_norm_module.num_features, batch_norm_module.eps, batch_norm_module.momentum,
                     batch_norm_module.affine, batch_norm_module.track_running_stats)
        module._adopt(batch_norm_module)
        if batch_norm_module.track_running_stats:
            module.running_mean = batch_norm_module.running_mean.clone()
            module.running_var = batch_norm_module.running_var.clone()
            module.num_batches_tracked = batch_norm_module.num_batches_tracked.clone()
        return module


class MaskedBatchNorm1d(_MaskedBatchNorm):
    def _check_input_dim(self, input: Tensor) -> None:
        if input.dim() not in (2, 3):
            raise ValueError(f"expected 2D or 3D input (got {input.dim()}D input)")


class MaskedBatchNorm2d(_MaskedBatchNorm):
    def _check_input_dim(self, input: Tensor) -> None:
        if input.dim() != 4:
            raise ValueError(f"expected 4D input (got {input.dim()}D input)")


class MaskedBatchNorm3d(_MaskedBatchNorm):
    def _check_input_dim(self, input: Tensor) -> None:
        if input.dim() != 5:
            raise ValueError(f"expected 5D input (got {input.dim()}D input)")
