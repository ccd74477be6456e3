"""MOON model: base -> (optional projection) -> head, features always flattened (parity: ``moon_base.py:7-45``)."""

from __future__ import annotations

import torch
from torch import nn

from fl4health_b200.model_bases.sequential_split_models import SequentiallySplitModel


class MoonModel(SequentiallySplitModel):
    def __init__(self, base_module: nn.Module, head_module: nn.Module, projection_module: nn.Module | None = None) -> None:
        super().__init__(base_module, head_module, flatten_features=True)
        self.projection_module = projection_module

    def sequential_forward(self, input: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        features = self.base_module(input)
        if self.projection_module is not None:
            features = self.projection_module(features)
        return self.head_module(features), features
