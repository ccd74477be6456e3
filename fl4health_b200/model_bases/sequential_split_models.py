"""``base_module -> head_module`` models (parity: ``fl4health/model_bases/sequential_split_models.py:7-107``).

Forward returns ``({"prediction": ...}, {"features": ...})``; the exchange variant shares only ``base_module.*``
(FedPer / FedRep / MOON building block).
"""

from __future__ import annotations

import torch
from torch import nn

from fl4health_b200.model_bases.partial_layer_exchange_model import PartialLayerExchangeModel


class SequentiallySplitModel(nn.Module):
    def __init__(self, base_module: nn.Module, head_module: nn.Module, flatten_features: bool = False) -> None:
        super().__init__()
        self.base_module = base_module
        self.head_module = head_module
        self.flatten_features = flatten_features

    def _flatten_features(self, features: torch.Tensor) -> torch.Tensor:
        return features.reshape(len(features), -1)

    def sequential_forward(self, input: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        features = self.base_module(input)
        return self.head_module(features), features

    def features_forward(self, input: torch.Tensor) -> torch.Tensor:
        features = self.base_module(input)
        return self._flatten_features(features) if self.flatten_features else features

    def forward(self, input: torch.Tensor) -> tuple[dict[str, torch.Tensor], dict[str, torch.Tensor]]:
        predictions, features = self.sequential_forward(input)
        if self.flatten_features:
            features = self._flatten_features(features)
        return {"prediction": predictions}, {"features": features}


class SequentiallySplitExchangeBaseModel(SequentiallySplitModel, PartialLayerExchangeModel):
    def layers_to_exchange(self) -> list[str]:
        return [name for name in self.state_dict() if name.startswith("base_module.")]
