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
_ATTRIBUTE = dict(zip(_ORDER, ("true_positives", "false_positives", "true_negatives", "false_negatives")))


def _outcome_products(p: torch.Tensor, t: torch.Tensor, wanted: tuple[int, ...]) -> torch.Tensor:
    """Soft outcome indicators stacked on a new leading axis, in ``_ORDER`` positions ``wanted``:
    TP = p t, FP = p (1 - t), TN = (1 - p)(1 - t), FN = (1 - p) t."""
    sides = {0: (p, t), 1: (p, 1 - t), 2: (1 - p, 1 - t), 3: (1 - p, t)}
    return torch.stack([sides[i][0] * sides[i][1] for i in wanted])


class _CountState:
    """The running counts of the outcomes that are kept, as ONE tensor ``[n_kept, ...]`` (one add / concatenate per
    update instead of four); the four per-outcome tensors of the public API are views of it."""

    def __init__(self, kept: tuple[int, ...], grows_along_batch: bool) -> None:
        self.kept, self.grows_along_batch = kept, grows_along_batch
        self.stacked: torch.Tensor | None = None

    def absorb(self, counts: Counts) -> None:
        if not self.kept:
            self.stacked = torch.empty(0)
            return
        fresh = torch.stack([counts[i] for i in self.kept])
        if self.stacked is None:
            self.stacked = fresh
        elif self.grows_along_batch:
            self.stacked = torch.cat([self.stacked, fresh], dim=1)
        else:
            self.stacked = self.stacked + fresh

    def of(self, outcome: int) -> torch.Tensor:
        if self.stacked is None or outcome not in self.kept:
            return torch.tensor([])
        return self.stacked[self.kept.index(outcome)]


class ClassificationMetric(Metric, ABC):
    def __init__(
        self, name: str, label_dim: int | None, batch_dim: int | None, dtype: torch.dtype,
        threshold: float | int | None, discard: set[ClassificationOutcome] | None,
    ) -> None:
        super().__init__(name)
        self.dtype, self.threshold, self.label_dim, self.batch_dim = dtype, threshold, label_dim, batch_dim
        if label_dim is not None and batch_dim is not None and label_dim == batch_dim:
            raise ValueError(f"The label and batch dimensions must differ but got {label_dim}")
        threshold_is_an_axis = isinstance(threshold, int) and not isinstance(threshold, bool)
        if label_dim is not None and threshold_is_an_axis and threshold != label_dim:
            log(WARNING, f"Specified threshold dimension: {threshold} is not the same as the label_dim: {label_dim}. "
                         "This is atypical and may produce undesired behavior")
        self.discarded = frozenset(discard or ())
        self.clear()

    # -- state ---------------------------------------------------------------------------------------------
    def clear(self) -> None:
        self._state = _CountState(self._kept_after_relabelling(), self.batch_dim is not None)

    def _kept_after_relabelling(self) -> tuple[int, ...]:
        """Positions (in ``_ORDER``) of the *returned* counts that are not empty."""
        return tuple(i for i, outcome in enumerate(_ORDER) if outcome not in self.discarded)

    @property
    def counts_initialized(self) -> bool:
        return self._state.stacked is not None

    true_positives = property(lambda self: self._state.of(0))
    false_positives = property(lambda self: self._state.of(1))
    true_negatives = property(lambda self: self._state.of(2))
    false_negatives = property(lambda self: self._state.of(3))
    discard_tp = property(lambda self: ClassificationOutcome.TRUE_POSITIVE in self.discarded)
    discard_fp = property(lambda self: ClassificationOutcome.FALSE_POSITIVE in self.discarded)
    discard_tn = property(lambda self: ClassificationOutcome.TRUE_NEGATIVE in self.discarded)
    discard_fn = property(lambda self: ClassificationOutcome.FALSE_NEGATIVE in self.discarded)

    # -- hooks for subclasses ------------------------------------------------------------------------------
    def _transform_tensors(self, preds: torch.Tensor, targets: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        preds, targets = (x.to(torch.uint8) if x.dtype == torch.bool else x for x in (preds, targets))
        if self.threshold is not None:
            preds = threshold_tensor(preds, self.threshold)
        return preds, targets

    def _assert_correct_ranges_and_shape(self, preds: torch.Tensor, targets: torch.Tensor) -> None:
        extremes = torch.stack([preds.min().float(), targets.min().float(), preds.max().float(), targets.max().float()])
        assert bool((extremes[:2].min() >= 0) & (extremes[2:].max() <= 1)), "Expected preds and targets to be in range [0, 1]."

    # -- counting ------------------------------------------------------------------------------------------
    def count_tp_fp_tn_fn(self, preds: torch.Tensor, targets: torch.Tensor) -> Counts:
        """Counts in the order (TP, FP, TN, FN); shapes: ``[]`` / ``[B]`` / ``[L]`` / ``[B, L]`` depending on which of
        batch_dim / label_dim are set (batch axis first).  Discarded outcomes come back as empty tensors."""
        preds, targets = self._transform_tensors(preds, targets)
        self._assert_correct_ranges_and_shape(preds, targets)
        wanted = tuple(i for i, outcome in enumerate(_ORDER) if outcome not in self.discarded)
        counts: list[torch.Tensor] = [torch.tensor([])] * len(_ORDER)
        if not wanted:
            return counts[0], counts[1], counts[2], counts[3]
        kept_axes = [axis for axis in (self.batch_dim, self.label_dim) if axis is not None]
        summed_axes = tuple(axis + 1 for axis in range(preds.ndim) if axis not in kept_axes)
        reduced = _outcome_products(preds.to(torch.float32), targets.to(torch.float32), wanted)
        if summed_axes:
            reduced = reduced.sum(summed_axes)
        reduced = reduced.to(self.dtype)
        if len(kept_axes) == MAX_COUNT_TENSOR_DIMS and kept_axes[0] > kept_axes[1]:  # batch axis goes first
            reduced = reduced.transpose(1, 2)
        for slot, position in enumerate(wanted):
            counts[position] = reduced[slot]
        return counts[0], counts[1], counts[2], counts[3]

    def update(self, preds: torch.Tensor, targets: torch.Tensor) -> None:
        self._state.absorb(self.count_tp_fp_tn_fn(preds, targets))

    def compute(self, name: str | None = None) -> Metrics:
        metrics = self.compute_from_counts(**{_ATTRIBUTE[outcome]: self._state.of(i) for i, outcome in enumerate(_ORDER)})
        return metrics if name is None else {f"{name} - {key}": value for key, value in metrics.items()}

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
        assert pos_label in {0, 1}, "pos_label must be either 0 or 1"
        self.pos_label = pos_label  # before the base constructor: it decides which returned counts are empty
        super().__init__(name=name, dtype=dtype, label_dim=label_dim, batch_dim=batch_dim, threshold=threshold, discard=discard)

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

    def _kept_after_relabelling(self) -> tuple[int, ...]:
        raw = super()._kept_after_relabelling()
        # with pos_label == 0 positives and negatives trade places in what is returned (raw TN is reported as TP, ...)
        return raw if getattr(self, "pos_label", 1) == 1 else tuple(sorted((i + 2) % 4 for i in raw))

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
