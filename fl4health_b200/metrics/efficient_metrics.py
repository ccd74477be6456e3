"""Dice metrics on streaming counts (parity: ``fl4health/metrics/efficient_metrics.py:15-307``)."""

from __future__ import annotations

from logging import WARNING

import torch

from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Metrics, Scalar
from fl4health_b200.metrics.efficient_metrics_base import (
    BinaryClassificationMetric,
    ClassificationOutcome,
    MultiClassificationMetric,
)
from fl4health_b200.metrics.metrics_utils import compute_dice_on_count_tensors


class _DiceMixin:
    name: str
    zero_division: float | None

    def compute_from_counts(self, true_positives: torch.Tensor, false_positives: torch.Tensor,
                            true_negatives: torch.Tensor, false_negatives: torch.Tensor) -> Metrics:  # noqa: ARG002
        dice = compute_dice_on_count_tensors(true_positives, false_positives, false_negatives, self.zero_division)
        if dice.numel() == 0:
            log(WARNING, "Currently, Dice score is undefined due to only true negatives present")
        return {self.name: torch.mean(dice).item()}

    def __call__(self, input: torch.Tensor, target: torch.Tensor) -> Scalar:
        counts = self.count_tp_fp_tn_fn(input, target)  # type: ignore[attr-defined]
        return self.compute_from_counts(*counts)[self.name]


class MultiClassDice(_DiceMixin, MultiClassificationMetric):
    """Mean Dice over labels (and over samples when ``batch_dim`` is given)."""

    def __init__(
        self, batch_dim: int | None, label_dim: int, name: str = "MultiClassDice", dtype: torch.dtype = torch.float32,
        threshold: float | int | None = None, ignore_background: int | None = None, zero_division: float | None = None,
    ) -> None:
        MultiClassificationMetric.__init__(
            self, name=name, batch_dim=batch_dim, label_dim=label_dim, dtype=dtype, threshold=threshold,
            ignore_background=ignore_background, discard={ClassificationOutcome.TRUE_NEGATIVE},
        )
        self.zero_division = zero_division


class BinaryDice(_DiceMixin, BinaryClassificationMetric):
    """Dice w.r.t. ``pos_label`` (mean over samples when ``batch_dim`` is given)."""

    def __init__(
        self, batch_dim: int | None, name: str = "BinaryDice", label_dim: int | None = None, dtype: torch.dtype = torch.float32,
        pos_label: int = 1, threshold: float | int | None = None, zero_division: float | None = None,
    ) -> None:
        # with pos_label == 0 the roles of (tp, fp, tn, fn) swap: the outcome Dice ignores is then the TRUE POSITIVES
        discard = {ClassificationOutcome.TRUE_NEGATIVE} if pos_label == 1 else {ClassificationOutcome.TRUE_POSITIVE}
        BinaryClassificationMetric.__init__(
            self, name=name, batch_dim=batch_dim, label_dim=label_dim, dtype=dtype, pos_label=pos_label,
            threshold=threshold, discard=discard,
        )
        self.zero_division = zero_division
