"""Two parallel feature extractors feeding one head (parity: ``parallel_split_models.py:8-124``)."""

from __future__ import annotations

from abc import ABC, abstractmethod
from enum import Enum

import torch
from torch import nn


class ParallelFeatureJoinMode(Enum):
    CONCATENATE = "CONCATENATE"
    SUM = "SUM"


class ParallelSplitHeadModule(nn.Module, ABC):
    def __init__(self, mode: ParallelFeatureJoinMode) -> None:
        super().__init__()
        self.mode = mode

    @abstractmethod
    def parallel_output_join(self, local_tensor: torch.Tensor, global_tensor: torch.Tensor) -> torch.Tensor:
        """How to concatenate the two feature tensors (only used in CONCATENATE mode)."""
        raise NotImplementedError

    @abstractmethod
    def head_forward(self, input_tensor: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def forward(self, first_tensor: torch.Tensor, second_tensor: torch.Tensor) -> torch.Tensor:
        if self.mode == ParallelFeatureJoinMode.CONCATENATE:
            joined = self.parallel_output_join(first_tensor, second_tensor)
        else:
            joined = first_tensor + second_tensor
        return self.head_forward(joined)


class ParallelSplitModel(nn.Module):
    def __init__(self, first_feature_extractor: nn.Module, second_feature_extractor: nn.Module,
                 model_head: ParallelSplitHeadModule) -> None:
        super().__init__()
        self.first_feature_extractor = first_feature_extractor
        self.second_feature_extractor = second_feature_extractor
        self.model_head = model_head

    def forward(self, input: torch.Tensor) -> tuple[dict[str, torch.Tensor], dict[str, torch.Tensor]]:
        first_output = self.first_feature_extractor(input)
        second_output = self.second_feature_extractor(input)
        return {"prediction": self.model_head(first_output, second_output)}, {}
