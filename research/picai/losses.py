"""Focal loss for binary lesion segmentation (parity: ``research/picai/losses.py:6-52``), computed from logits in one
numerically stable expression instead of ``sigmoid`` followed by ``binary_cross_entropy``."""

from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn


class FocalLoss(nn.Module):
    """``alpha_t * (1 - p_t)^gamma * CE`` with ``alpha`` the positive-class weight (``alpha < 0`` disables it)."""

    def __init__(self, alpha: float = 1.0, gamma: float = 1.0, reduction: str = "sum") -> None:
        super().__init__()
        if reduction not in ("sum", "mean"):
            raise NotImplementedError(f"reduction '{reduction}'")
        self.alpha, self.gamma, self.reduction = alpha, gamma, reduction

    def forward(self, inputs: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        targets = targets.to(inputs.dtype)
        ce = F.binary_cross_entropy_with_logits(inputs, targets, reduction="none")
        p_t = torch.exp(-ce)  # probability assigned to the true class
        loss = ce * (1.0 - p_t) ** self.gamma
        if self.alpha >= 0:
            loss = (self.alpha * targets + (1.0 - self.alpha) * (1.0 - targets)) * loss
        return loss.mean() if self.reduction == "mean" else loss.sum()
