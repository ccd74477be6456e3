"""Parity: ``fl4health/metrics/metrics_utils.py:4-81``."""

from __future__ import annotations

import torch


def compute_dice_on_count_tensors(
    true_positives: torch.Tensor, false_positives: torch.Tensor, false_negatives: torch.Tensor, zero_division: float | None
) -> torch.Tensor:
    """Elementwise ``2 TP / (2 TP + FP + FN)``, flattened.  Undefined entries (denominator 0) are dropped when
    ``zero_division`` is None and replaced by it otherwise."""
    numerator = (2 * true_positives).reshape(-1).to(torch.float32)
    denominator = (2 * true_positives + false_positives + false_negatives).reshape(-1).to(torch.float32)
    undefined = denominator == 0
    if zero_division is None:
        keep = ~undefined
        return numerator[keep] / denominator[keep]
    return torch.where(undefined, torch.full_like(numerator, float(zero_division)), numerator / denominator.clamp_min(1e-38))


def threshold_tensor(input: torch.Tensor, threshold: float | int) -> torch.Tensor:
    """float: ``x > threshold``; int: one-hot of the arg-max along that axis."""
    if isinstance(threshold, bool) or not isinstance(threshold, (float, int)):
        raise ValueError(f"Was expecting threshold argument to be either a float or an int. Got {type(threshold)}")
    if isinstance(threshold, float):
        return (input > threshold).to(input.dtype)
    if threshold >= input.ndim:
        raise ValueError(
            f"Cannot apply argmax to Tensor of shape {input.shape}. Label dimension of {threshold} is out of range of "
            f"tensor with {input.ndim} dimensions."
        )
    return torch.zeros_like(input).scatter_(threshold, input.argmax(threshold, keepdim=True), 1)
