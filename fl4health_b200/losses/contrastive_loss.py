"""Contrastive losses (parity: ``fl4health/losses/contrastive_loss.py:6-167``).

``MoonContrastiveLoss``: pull features towards one positive view and away from N negative views (MOON, PerFCL,
constrained FENDA) — a single fused kernel forward/backward on CUDA.  ``NtXentLoss``: SimCLR's normalised-temperature
cross entropy over the 2B x 2B similarity matrix.
"""

from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from fl4health_b200.ops.contrastive import moon_contrastive


class MoonContrastiveLoss(nn.Module):
    def __init__(self, device: torch.device, temperature: float = 0.5) -> None:
        super().__init__()
        self.device = device
        self.temperature = temperature

    def compute_negative_similarities(self, features: torch.Tensor, negative_pairs: torch.Tensor) -> torch.Tensor:
        """cos-sim between ``features`` [B,F] and each of ``negative_pairs`` [N,B,F] -> [N,B]."""
        assert features.shape == negative_pairs.shape[1:]
        return F.cosine_similarity(features.unsqueeze(0), negative_pairs, dim=-1)

    def forward(self, features: torch.Tensor, positive_pairs: torch.Tensor, negative_pairs: torch.Tensor) -> torch.Tensor:
        features = features.to(self.device)
        positive_pairs, negative_pairs = positive_pairs.to(self.device), negative_pairs.to(self.device)
        if len(positive_pairs) != 1:
            raise AssertionError(
                "Each feature can have only one positive pair. Thus positive pairs should be a tensor of shape "
                f"(1, batch_size, n_features) rather than {positive_pairs.shape}"
            )
        positive = positive_pairs[0]
        assert len(features) == len(positive)
        assert features.shape == negative_pairs.shape[1:]
        return moon_contrastive(features, positive, negative_pairs, self.temperature)


class NtXentLoss(nn.Module):
    def __init__(self, device: torch.device, temperature: float = 0.5) -> None:
        super().__init__()
        self.device = device
        self.temperature = temperature

    def forward(self, features: torch.Tensor, transformed_features: torch.Tensor) -> torch.Tensor:
        assert features.shape == transformed_features.shape
        batch = features.shape[0]
        both = F.normalize(torch.cat([features, transformed_features], dim=0).to(self.device), dim=-1)
        similarity = both @ both.t()
        positives = torch.cat([torch.diag(similarity, batch), torch.diag(similarity, -batch)], dim=0)
        # the reference zeroes the diagonal *similarity* (so exp(0)=1 stays in the denominator); keep that convention
        off_diagonal = similarity * (1.0 - torch.eye(2 * batch, device=both.device, dtype=both.dtype))
        denominator = torch.exp(off_diagonal / self.temperature).sum(dim=1)
        losses = -(positives / self.temperature) + torch.log(denominator)
        return losses.sum() / (2 * batch)
