"""Weighted loss bundles for constrained FENDA (parity: ``fl4health/losses/fenda_loss_config.py:8-157``).

Each optional regulariser is a *weighted term*: a weight and a loss module.  The three public containers keep the
reference's attribute names (clients and configs read them), but they are thin views over ``_WeightedTerm`` and the
aggregate container computes ``weight * loss(...)`` through one code path."""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any

import torch

from fl4health_b200.losses.contrastive_loss import MoonContrastiveLoss
from fl4health_b200.losses.cosine_similarity_loss import CosineSimilarityLoss
from fl4health_b200.losses.perfcl_loss import PerFclLoss


@dataclass
class _WeightedTerm:
    weights: tuple[float, ...]
    function: Any

    def __call__(self, *tensors: torch.Tensor) -> Any:
        value = self.function(*tensors)
        if isinstance(value, tuple):  # multi-part losses carry one weight per part
            return tuple(weight * part for weight, part in zip(self.weights, value))
        return self.weights[0] * value


class PerFclLossContainer:
    def __init__(self, device: torch.device, global_feature_contrastive_loss_weight: float,
                 local_feature_contrastive_loss_weight: float, global_feature_loss_temperature: float = 0.5,
                 local_feature_loss_temperature: float = 0.5) -> None:
        self.global_feature_contrastive_loss_weight = global_feature_contrastive_loss_weight
        self.local_feature_contrastive_loss_weight = local_feature_contrastive_loss_weight
        self.perfcl_loss_function = PerFclLoss(device, global_feature_loss_temperature, local_feature_loss_temperature)

    def term(self) -> _WeightedTerm:
        return _WeightedTerm((self.global_feature_contrastive_loss_weight, self.local_feature_contrastive_loss_weight),
                             self.perfcl_loss_function)


class CosineSimilarityLossContainer:
    def __init__(self, device: torch.device, cos_sim_loss_weight: float) -> None:
        self.cos_sim_loss_weight = cos_sim_loss_weight
        self.cos_sim_loss_function = CosineSimilarityLoss(device)

    def term(self) -> _WeightedTerm:
        return _WeightedTerm((self.cos_sim_loss_weight,), self.cos_sim_loss_function)


class MoonContrastiveLossContainer:
    def __init__(self, device: torch.device, contrastive_loss_weight: float, temperature: float = 0.5) -> None:
        self.contrastive_loss_weight = contrastive_loss_weight
        self.contrastive_loss_function = MoonContrastiveLoss(device, temperature)

    def term(self) -> _WeightedTerm:
        return _WeightedTerm((self.contrastive_loss_weight,), self.contrastive_loss_function)


class ConstrainedFendaLossContainer:
    def __init__(self, perfcl_loss_config: PerFclLossContainer | None,
                 cosine_similarity_loss_config: CosineSimilarityLossContainer | None,
                 contrastive_loss_config: MoonContrastiveLossContainer | None) -> None:
        self.perfcl_loss_config = perfcl_loss_config
        self.cos_sim_loss_config = cosine_similarity_loss_config
        self.contrastive_loss_config = contrastive_loss_config

    def _term(self, which: str) -> _WeightedTerm:
        config = getattr(self, which)
        assert config is not None, f"{which} was not configured"
        return config.term()

    def has_perfcl_loss(self) -> bool:
        return self.perfcl_loss_config is not None

    def has_cosine_similarity_loss(self) -> bool:
        return self.cos_sim_loss_config is not None

    def has_contrastive_loss(self) -> bool:
        return self.contrastive_loss_config is not None

    def compute_contrastive_loss(self, features: torch.Tensor, positive_pairs: torch.Tensor, negative_pairs: torch.Tensor) -> torch.Tensor:
        return self._term("contrastive_loss_config")(features, positive_pairs, negative_pairs)

    def compute_cosine_similarity_loss(self, first_features: torch.Tensor, second_features: torch.Tensor) -> torch.Tensor:
        return self._term("cos_sim_loss_config")(first_features, second_features)

    def compute_perfcl_loss(
        self, local_features: torch.Tensor, old_local_features: torch.Tensor, global_features: torch.Tensor,
        old_global_features: torch.Tensor, initial_global_features: torch.Tensor,
    ) -> tuple[torch.Tensor, torch.Tensor]:
        return self._term("perfcl_loss_config")(local_features, old_local_features, global_features, old_global_features,
                                                initial_global_features)
