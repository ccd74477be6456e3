"""FedSimCLR model (parity: ``fl4health/model_bases/fedsimclr_base.py:12-85``): encoder + projection head for
contrastive pre-training, encoder + prediction head for fine-tuning."""

from __future__ import annotations

from pathlib import Path

import torch
from torch import nn

DEFAULT_PROJECTION_HEAD = nn.Identity()


class FedSimClrModel(nn.Module):
    def __init__(
        self, encoder: nn.Module, projection_head: nn.Module | None = None, prediction_head: nn.Module | None = None,
        pretrain: bool = True,
    ) -> None:
        super().__init__()
        assert not (prediction_head is None and not pretrain), "Model with pretrain==False must have prediction head (ie not None)"
        self.encoder = encoder
        self.projection_head = projection_head if projection_head is not None else nn.Identity()
        self.prediction_head = prediction_head
        self.pretrain = pretrain

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        features = self.encoder(input)
        if self.pretrain:
            return self.projection_head(features)
        assert self.prediction_head is not None, "Model with pretrain==False must have prediction_head (ie not None)"
        return self.prediction_head(features)

    @staticmethod
    def load_pretrained_model(model_path: Path) -> FedSimClrModel:
        previous = torch.load(model_path, weights_only=False)
        return FedSimClrModel(previous.encoder, previous.projection_head, previous.prediction_head, pretrain=False)
