"""Mean absolute cosine similarity (parity: ``fl4health/losses/cosine_similarity_loss.py:5-30``)."""

from __future__ import annotations

import torch
from torch import nn


class CosineSimilarityLoss(nn.Module):
    def __init__(self, device: torch.device, dim: int = -1) -> None:
        super().__init__()
        self.dim = dim
        self.device = device

    def forward(self, x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
        assert len(x1) == len(x2), "Tensors have different batch sizes"
        return torch.nn.functional.cosine_similarity(x1, x2, dim=self.dim).abs().mean()
