"""Ensembles of models with average / majority-vote aggregation (parity: ``ensemble_base.py:15-107``).
The vote is a vectorised ``one_hot`` + ``sum`` + ``argmax`` instead of a Python loop over samples (:81-86)."""

from __future__ import annotations

from enum import Enum

import torch
from torch import nn

EXPECTED_MAX_PRED_N_DIMS = 2  # [batch, classes] predictions per ensemble member


class EnsembleAggregationMode(Enum):
    VOTE = "VOTE"
    AVERAGE = "AVERAGE"


class EnsembleModel(nn.Module):
    def __init__(self, ensemble_models: dict[str, nn.Module],
                 aggregation_mode: EnsembleAggregationMode | None = EnsembleAggregationMode.AVERAGE) -> None:
        super().__init__()
        self.ensemble_models = nn.ModuleDict(ensemble_models)
        self.aggregation_mode = aggregation_mode

    def forward(self, input: torch.Tensor) -> dict[str, torch.Tensor]:
        preds = {key: model(input) for key, model in self.ensemble_models.items()}
        stacked = list(preds.values())
        if self.aggregation_mode == EnsembleAggregationMode.AVERAGE:
            preds["ensemble-pred"] = self.ensemble_average(stacked)
        else:
            preds["ensemble-pred"] = self.ensemble_vote(stacked)
        return preds

    def ensemble_vote(self, preds_list: list[torch.Tensor]) -> torch.Tensor:
        """One-hot of the per-sample majority class (ties -> lowest class index)."""
        shape = preds_list[0].shape
        n_classes = shape[-1]
        votes = torch.stack([p.reshape(-1, n_classes).argmax(dim=1) for p in preds_list])  # [models, samples]
        counts = torch.nn.functional.one_hot(votes, n_classes).sum(dim=0)  # [samples, classes]
        winners = counts.argmax(dim=1)
        return torch.nn.functional.one_hot(winners, n_classes).reshape(shape).to(preds_list[0].dtype)

    def ensemble_average(self, preds_list: list[torch.Tensor]) -> torch.Tensor:
        return torch.stack(preds_list).mean(dim=0)
