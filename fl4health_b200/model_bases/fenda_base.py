"""FENDA: local + global parallel extractors; only the global (second) extractor is exchanged
(parity: ``fenda_base.py:8-80``)."""

from __future__ import annotations

import torch
from torch import nn

from fl4health_b200.model_bases.parallel_split_models import ParallelSplitHeadModule, ParallelSplitModel
from fl4health_b200.model_bases.partial_layer_exchange_model import PartialLayerExchangeModel


class FendaModel(ParallelSplitModel, PartialLayerExchangeModel):
    def __init__(self, local_module: nn.Module, global_module: nn.Module, model_head: ParallelSplitHeadModule) -> None:
        super().__init__(first_feature_extractor=local_module, second_feature_extractor=global_module, model_head=model_head)

    def layers_to_exchange(self) -> list[str]:
        return [name for name in self.state_dict() if name.startswith("second_feature_extractor.")]


class FendaModelWithFeatureState(FendaModel):
    """Also returns the two (optionally flattened) feature tensors for constraint losses."""

    def __init__(self, local_module: nn.Module, global_module: nn.Module, model_head: ParallelSplitHeadModule,
                 flatten_features: bool = False) -> None:
        super().__init__(local_module, global_module, model_head)
        self.flatten_features = flatten_features

    def forward(self, input: torch.Tensor) -> tuple[dict[str, torch.Tensor], dict[str, torch.Tensor]]:
        local_output = self.first_feature_extractor(input)
        global_output = self.second_feature_extractor(input)
        preds = {"prediction": self.model_head(local_output, global_output)}
        if self.flatten_features:
            local_output, global_output = local_output.reshape(len(local_output), -1), global_output.reshape(len(global_output), -1)
        return preds, {"local_features": local_output, "global_features": global_output}
