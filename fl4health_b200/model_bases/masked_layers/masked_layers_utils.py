"""Model surgery for FedPM (parity: ``masked_layers_utils.py:23-96``)."""

from __future__ import annotations

import copy

from torch import nn

from fl4health_b200.model_bases.masked_layers.masked_layers import (
    MaskedBatchNorm1d,
    MaskedBatchNorm2d,
    MaskedBatchNorm3d,
    MaskedConv1d,
    MaskedConv2d,
    MaskedConv3d,
    MaskedConvTranspose1d,
    MaskedConvTranspose2d,
    MaskedConvTranspose3d,
    MaskedLayerNorm,
    MaskedLinear,
    _MaskedBatchNorm,
)

_REPLACEMENTS: list[tuple[type, type]] = [
    (nn.Linear, MaskedLinear), (nn.Conv1d, MaskedConv1d), (nn.Conv2d, MaskedConv2d), (nn.Conv3d, MaskedConv3d),
    (nn.ConvTranspose1d, MaskedConvTranspose1d), (nn.ConvTranspose2d, MaskedConvTranspose2d),
    (nn.ConvTranspose3d, MaskedConvTranspose3d), (nn.LayerNorm, MaskedLayerNorm),
    (nn.BatchNorm1d, MaskedBatchNorm1d), (nn.BatchNorm2d, MaskedBatchNorm2d), (nn.BatchNorm3d, MaskedBatchNorm3d),
]
_MASKED_TYPES = tuple(masked for _, masked in _REPLACEMENTS) + (_MaskedBatchNorm,)


def is_masked_module(module: nn.Module) -> bool:
    return isinstance(module, _MASKED_TYPES)


def convert_to_masked_model(original_model: nn.Module) -> nn.Module:
    """Deep-copied model in which every supported layer is replaced by its masked counterpart (recursively)."""

    def replace(module: nn.Module) -> None:
        for name, child in module.named_children():
            if is_masked_module(child):
                continue
            for stock, masked in _REPLACEMENTS:
                if type(child) is stock:
                    setattr(module, name, masked.from_pretrained(child))
                    break
            else:
                replace(child)

    masked_model = copy.deepcopy(original_model)
    if not is_masked_module(masked_model):
        for stock, masked in _REPLACEMENTS:
            if type(masked_model) is stock:
                return masked.from_pretrained(masked_model)
    replace(masked_model)
    return masked_model
