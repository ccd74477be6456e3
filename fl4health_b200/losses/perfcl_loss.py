"""PerFCL's two contrastive terms (parity: ``fl4health/losses/perfcl_loss.py:7-91``):
global extractor: stay close to the aggregated extractor's features, away from last round's own;
local extractor:  stay close to last round's local features, away from the aggregated global ones."""

from __future__ import annotations

import torch
from torch import nn

from fl4health_b200.losses.contrastive_loss import MoonContrastiveLoss


class PerFclLoss(nn.Module):
    def __init__(self, device: torch.device, global_feature_loss_temperature: float = 0.5,
                 local_feature_loss_temperature: float = 0.5) -> None:
        super().__init__()
        self.global_feature_contrastive_loss = MoonContrastiveLoss(device, global_feature_loss_temperature)
        self.local_feature_contrastive_loss = MoonContrastiveLoss(device, local_feature_loss_temperature)

    def forward(
        self, local_features: torch.Tensor, old_local_features: torch.Tensor, global_features: torch.Tensor,
        old_global_features: torch.Tensor, initial_global_features: torch.Tensor,
    ) -> tuple[torch.Tensor, torch.Tensor]:
        z_g = initial_global_features.unsqueeze(0)
        global_loss = self.global_feature_contrastive_loss(
            features=global_features, positive_pairs=z_g, negative_pairs=old_global_features.unsqueeze(0)
        )
        local_loss = self.local_feature_contrastive_loss(
            features=local_features, positive_pairs=old_local_features.unsqueeze(0), negative_pairs=z_g
        )
        return global_loss, local_loss
