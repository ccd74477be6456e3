"""PerFCL model: FENDA layout whose features are always flattened (parity: ``perfcl_base.py:8-58``)."""

from __future__ import annotations

import torch
from torch import nn

from fl4health_b200.model_bases.parallel_split_models import ParallelSplitHeadModule, ParallelSplitModel
from fl4health_b200.model_bases.partial_layer_exchange_model import PartialLayerExchangeModel


class PerFclModel(ParallelSplitModel, PartialLayerExchangeModel):
    def __init__(self, local_module: nn.Module, global_module: nn.Module, model_head: ParallelSplitHeadModule) -> None:
        super().__init__(first_feature_extractor=local_module, second_feature_extractor=global_module, model_head=model_head)

    def layers_to_exchange(self) -> list[str]:
        return [name for name in self.state_dict() if name.startswith("second_feature_extractor.")]

    def forward(self, input: torch.Tensor) -> tuple[dict[str, torch.Tensor], dict[str, torch.Tensor]]:
        local_output = self.first_feature_extractor(input)
        global_output = self.second_feature_extractor(input)
        preds = {"prediction": self.model_head(local_output, global_output)}
        return preds, {
            "local_features": local_output.reshape(len(local_output), -1),
            "global_features": global_output.reshape(len(global_output), -1),
        }
