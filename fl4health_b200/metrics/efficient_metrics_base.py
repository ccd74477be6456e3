"""Streaming confusion-count metrics (parity: ``fl4health/metrics/efficient_metrics_base.py:18-920``).

State is four count tensors (TP / FP / TN / FN) reduced over every axis except the optional batch and label axes;
"soft" predictions in [0, 1] give soft counts.  Counts live on the device of the inputs and are only read by
``compute``.  All four outcomes of an update are produced from ONE stacked reduction (``stack -> sum``) instead of
four products and four reductions.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from enum import Enum
from logging import INFO, WARNING

import torch

from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Metrics, Scalar
from fl4health_b200.metrics.base_metrics import Metric
from fl4health_b200.metrics.metrics_utils import threshold_tensor
from fl4health_b200.metrics.utils import align_pred_and_target_shapes

MAX_COUNT_TENSOR_DIMS = 2  # count tensors are never more than 2-dimensional

N_LABELS_BINARY = 2


class ClassificationOutcome(Enum):
    TRUE_POSITIVE = "true_positive"
    FALSE_POSITIVE = "false_positive"
    TRUE_NEGATIVE = "true_negative"
    FALSE_NEGATIVE = "false_negative"


MetricOutcome = ClassificationOutcome  # the reference's older name for the same enum


_ORDER = (
    ClassificationOutcome.TRUE_POSITIVE, ClassificationOutcome.FALSE_POSITIVE,
    ClassificationOutcome.TRUE_NEGATIVE, ClassificationOutcome.FALSE_NEGATIVE,
)
Counts = tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]


class ClassificationMetric(Metric, ABC):
    def __init__(
        self, name: str, label_dim: int | None, batch_dim: int | None, dtype: torch.dtype,
        threshold: float | int | None, discard: set[ClassificationOutcome] | None,
    ) -> None:
        super().__init__(name)
        self.dtype, self.threshold, self.label_dim, self.batch_dim = dtype, threshold, label_dim, batch_dim
        if label_dim is not None:
            if isinstance(threshold, int) and not isinstance(threshold, bool) and threshold != label_dim:
                log(WARNING, f"Specified threshold dimension: {threshold} is not the same as the label_dim: {label_dim}. "
                             "This is atypical and may produce undesired behavior")
            if batch_dim is not None and label_dim == batch_dim:
                raise ValueError(f"The label and batch dimensions must differ but got {label_dim}")
        discard = discard or set()
        self.discard_tp = ClassificationOutcome.TRUE_POSITIVE in discard
        self.discard_fp = ClassificationOutcome.FALSE_POSITIVE in discard
        self.discard_tn = ClassificationOutcome.TRUE_NEGATIVE in discard
        self.discard_fn = ClassificationOutcome.FALSE_NEGATIVE in discard
        self.clear()

    # -- state ---------------------------------------------------------------------------------------------
    def clear(self) -> None:
        self.true_positives, self.false_positives = torch.tensor([]), torch.tensor([])
        self.true_negatives, self.false_negatives = torch.tensor([]), torch.tensor([])
        self.counts_initialized = False

    def _discarded(self) -> tuple[bool, bool, bool, bool]:
        return self.discard_tp, self.discard_fp, self.discard_tn, self.discard_fn

    # -- hooks for subclasses ------------------------------------------------------------------------------
    def _transform_tensors(self, preds: torch.Tensor, targets: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        preds = preds.to(torch.uint8) if preds.dtype == torch.bool else preds
        targets = targets.to(torch.uint8) if targets.dtype == torch.bool else targets
        if self.threshold is not None:
            preds = threshold_tensor(preds, self.threshold)
        return preds, targets

    def _assert_correct_ranges_and_shape(self, preds: torch.Tensor, targets: torch.Tensor) -> None:
        lo = torch.minimum(preds.min().float(), targets.min().float())
        hi = torch.maximum(preds.max().float(), targets.max().float())
        assert bool((lo >= 0) & (hi <= 1)), "Expected preds and targets to be in range [0, 1]."

    # -- counting ------------------------------------------------------------------------------------------
    def count_tp_fp_tn_fn(self, preds: torch.Tensor, targets: torch.Tensor) -> Counts:
        """Counts in the order (TP, FP, TN, FN); shapes: ``[]`` / ``[B]`` / ``[L]`` / ``[B, L]`` depending on which of
        batch_dim / label_dim are set (batch axis first).  Discarded outcomes come back as empty tensors."""
        preds, targets = self._transform_tensors(preds, targets)
        self._assert_correct_ranges_and_shape(preds, targets)
        keep_axes = {d for d in (self.label_dim, self.batch_dim) if d is not None}
        sum_axes = tuple(i for i in range(preds.ndim) if i not in keep_axes)
        p, t = preds.to(torch.float32), targets.to(torch.float32)
        # one stacked reduction for all live outcomes
        products = {0: lambda: p * t, 1: lambda: p * (1 - t), 2: lambda: (1 - p) * (1 - t), 3: lambda: (1 - p) * t}
        live = [i for i, dropped in enumerate(self._discarded()) if not dropped]
        out: list[torch.Tensor] = [torch.tensor([])] * 4
        if live:
            stacked = torch.stack([products[i]() for i in live])
            if sum_axes:
                stacked = stacked.sum(tuple(a + 1 for a in sum_axes))
            stacked = stacked.to(self.dtype)
            if stacked.ndim == 3 and self.batch_dim is not None and self.label_dim is not None and self.batch_dim > self.label_dim:
                stacked = stacked.transpose(1, 2)
            for slot, i in enumerate(live):
                out[i] = stacked[slot]
        return out[0], out[1], out[2], out[3]

    def update(self, preds: torch.Tensor, targets: torch.Tensor) -> None:
        tp, fp, tn, fn = self.count_tp_fp_tn_fn(preds, targets)
        if not self.counts_initialized:
            self.true_positives, self.false_positives, self.true_negatives, self.false_negatives = tp, fp, tn, fn
            self.counts_initialized = True
            return
        merge = (lambda a, b: torch.cat([a, b], dim=0)) if self.batch_dim is not None else (lambda a, b: a + b)
        self.true_positives = merge(self.true_positives, tp)
        self.false_positives = merge(self.false_positives, fp)
        self.true_negatives = merge(self.true_negatives, tn)
        self.false_negatives = merge(self.false_negatives, fn)

    def compute(self, name: str | None = None) -> Metrics:
        metrics = self.compute_from_counts(
            true_positives=self.true_positives, false_positives=self.false_positives,
            true_negatives=self.true_negatives, false_negatives=self.false_negatives,
        )
        return {f"{name} - {k}": v for k, v in metrics.items()} if name is not None else metrics

    @abstractmethod
    def compute_from_counts(self, true_positives: torch.Tensor, false_positives: torch.Tensor,
                            true_negatives: torch.Tensor, false_negatives: torch.Tensor) -> Metrics:
        raise NotImplementedError

    def __call__(self, input: torch.Tensor, target: torch.Tensor) -> Scalar:
        raise NotImplementedError


class BinaryClassificationMetric(ClassificationMetric):
    """Counts relative to ``pos_label`` only; label axis (if any) has at most 2 entries.  Count shapes are ``(1,)``
    or ``(num_samples, 1)`` when ``batch_dim`` is set."""

    def __init__(
        self, name: str, label_dim: int | None = None, batch_dim: int | None = None, dtype: torch.dtype = torch.float32,
        pos_label: int = 1, threshold: float | int | None = None, discard: set[ClassificationOutcome] | None = None,
    ) -> None:
        super().__init__(name=name, dtype=dtype, label_dim=label_dim, batch_dim=batch_dim, threshold=threshold, discard=discard)
        assert pos_label in {0, 1}, "pos_label must be either 0 or 1"
        self.pos_label = pos_label

    def _postprocess_count_tensor(self, count_tensor: torch.Tensor) -> torch.Tensor:
        if count_tensor.numel() == 0:
            return count_tensor
        if self.batch_dim is not None and self.label_dim is not None:
            assert count_tensor.ndim == 2, f"Batch and label dims have been specified, tensor should be 2D, but got {count_tensor.ndim}"
            if count_tensor.shape[1] == N_LABELS_BINARY:
                return count_tensor[:, 1:2]
            if count_tensor.shape[1] == 1:
                return count_tensor
            raise ValueError(f"Label dimension has unexpected size of {count_tensor.shape[1]}")
        if self.batch_dim is not None:
            assert count_tensor.ndim == 1, f"Batch dim has been specified but not label dim, tensor should be 1D but got {count_tensor.ndim}D"
            return count_tensor.unsqueeze(1)
        assert count_tensor.ndim <= 1, f"Batch dim has not been specified, tensor should be 0 or 1D but got {count_tensor.ndim}"
        if count_tensor.numel() == N_LABELS_BINARY:
            assert self.label_dim is not None, "self.label_dim is None but got two elements in the count_tensor"
            return count_tensor[1:2]
        if count_tensor.numel() == 1:
            return count_tensor.reshape(1)
        raise ValueError(f"Too many elements in the count tensor, expected 2 or less and got {count_tensor.numel()}")

    def _assert_correct_ranges_and_shape(self, preds: torch.Tensor, targets: torch.Tensor) -> None:
        super()._assert_correct_ranges_and_shape(preds, targets)
        assert preds.shape == targets.shape, (
            f"Preds and targets must have the same shape but got {preds.shape} and {targets.shape} respectively."
        )
        if self.label_dim is not None:
            for kind, tensor in (("preds", preds), ("targets", targets)):
                assert tensor.shape[self.label_dim] <= N_LABELS_BINARY, (
                    f"Label dimension for {kind} tensor is greater than 2 {tensor.shape[self.label_dim]}. This class is "
                    "meant for binary metric computation only"
                )

    def count_tp_fp_tn_fn(self, preds: torch.Tensor, targets: torch.Tensor) -> Counts:
        tp, fp, tn, fn = (self._postprocess_count_tensor(c) for c in super().count_tp_fp_tn_fn(preds, targets))
        return (tn, fn, tp, fp) if self.pos_label == 0 else (tp, fp, tn, fn)


class MultiClassificationMetric(ClassificationMetric):
    """Per-label counts (label axis of size >= 2); one side may be label-index encoded and is one-hot expanded."""

    def __init__(
        self, name: str, label_dim: int, batch_dim: int | None = None, dtype: torch.dtype = torch.float32,
        threshold: float | int | None = None, ignore_background: int | None = None,
        discard: set[ClassificationOutcome] | None = None,
    ) -> None:
        super().__init__(name=name, dtype=dtype, label_dim=label_dim, batch_dim=batch_dim, threshold=threshold, discard=discard)
        if ignore_background is not None:
            log(INFO, f"ignore_background has been specified. The first channel of dimension {ignore_background} "
                      "will be removed from both predictions and targets")
        self.ignore_background = ignore_background

    @classmethod
    def _remove_background(cls, ignore_background: int, preds: torch.Tensor, targets: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        assert preds.shape == targets.shape, f"Preds ({preds.shape}) and targets ({targets.shape}) should have the same shape but do not."
        size = preds.shape[ignore_background]
        return preds.narrow(ignore_background, 1, size - 1), targets.narrow(ignore_background, 1, size - 1)

    def _transform_tensors(self, preds: torch.Tensor, targets: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        preds, targets = super()._transform_tensors(preds, targets)
        preds, targets = align_pred_and_target_shapes(preds, targets, self.label_dim)
        if self.ignore_background is not None:
            preds, targets = self._remove_background(self.ignore_background, preds, targets)
        return preds, targets

    def _assert_correct_ranges_and_shape(self, preds: torch.Tensor, targets: torch.Tensor) -> None:
        super()._assert_correct_ranges_and_shape(preds, targets)
        assert self.label_dim is not None
        for kind, tensor in (("preds", preds), ("targets", targets)):
            assert tensor.shape[self.label_dim] >= N_LABELS_BINARY, (
                f"Label dimension for {kind} tensor is less than 2. Either your label dimension is a single float value "
                "corresponding to a binary prediction or it is a class label that needs to be vector encoded."
            )
