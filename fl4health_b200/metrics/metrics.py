"""Classification / segmentation metrics.

Parity: ``fl4health/metrics/metrics.py:12-247`` (``TorchMetric``, ``SimpleMetric``, ``Accuracy``,
``BalancedAccuracy``, ``RocAuc``, ``F1``, ``BinarySoftDiceCoefficient``).  The reference appends every batch's
logits to a Python list and runs sklearn on the CPU at ``compute`` (SURVEY hot-op L15).  Here accuracy, balanced
accuracy and F1 are *streaming on-device counters* (a correct/total pair or a confusion matrix) whose ``update`` is a
few capturable tensor ops; only rank statistics (ROC-AUC) and user ``SimpleMetric`` subclasses still buffer batches.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from collections.abc import Sequence
from typing import Any

import numpy as np
import torch

from fl4health_b200.common.typing import Metrics, Scalar
from fl4health_b200.metrics.base_metrics import Metric


class TorchMetric(Metric):
    """Adapter over any object with the torchmetrics protocol (``update/compute/reset``)."""

    def __init__(self, name: str, metric: Any) -> None:
        super().__init__(name)
        self.metric = metric

    def update(self, input: torch.Tensor, target: torch.Tensor) -> None:
        self.metric.update(input, target.long())

    def compute(self, name: str | None = None) -> Metrics:
        return {self._key(name): float(self.metric.compute().item())}

    def clear(self) -> None:
        self.metric.reset()


class SimpleMetric(Metric, ABC):
    """Buffers every batch and evaluates ``__call__`` over the concatenation at ``compute``."""

    def __init__(self, name: str) -> None:
        super().__init__(name)
        self.accumulated_inputs: list[torch.Tensor] = []
        self.accumulated_targets: list[torch.Tensor] = []

    def update(self, input: torch.Tensor, target: torch.Tensor) -> None:
        self.accumulated_inputs.append(input.detach())
        self.accumulated_targets.append(target.detach())

    def compute(self, name: str | None = None) -> Metrics:
        assert len(self.accumulated_inputs) > 0 and len(self.accumulated_targets) > 0
        result = self(torch.cat(self.accumulated_inputs), torch.cat(self.accumulated_targets))
        return {self._key(name): result}

    def clear(self) -> None:
        self.accumulated_inputs = []
        self.accumulated_targets = []

    @abstractmethod
    def __call__(self, input: torch.Tensor, target: torch.Tensor) -> Scalar:
        raise NotImplementedError


def _hard_predictions(logits: torch.Tensor, threshold: float = 0.5) -> torch.Tensor:
    if logits.dim() == 1 or logits.shape[1] == 1:
        return (logits.reshape(-1) > threshold).long()
    return torch.argmax(logits, dim=1)


def _hard_targets(logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    # one-hot / soft targets with the same shape as multi-class logits are reduced to class indices
    if target.dim() > 1 and target.shape == logits.shape and logits.dim() > 1 and logits.shape[1] > 1:
        return torch.argmax(target, dim=1)
    return target.reshape(-1).long()


class Accuracy(Metric):
    """Streaming accuracy: two device scalars, no host traffic until ``compute``."""

    def __init__(self, name: str = "accuracy") -> None:
        super().__init__(name)
        self.correct: torch.Tensor | None = None
        self.total: torch.Tensor | None = None

    def update(self, input: torch.Tensor, target: torch.Tensor) -> None:
        logits = input.detach()
        assert logits.shape[0] == target.shape[0], "Batch size of preds and targets do not match"
        preds = _hard_predictions(logits)
        hits = (preds.reshape(-1) == _hard_targets(logits, target)).sum()
        if self.correct is None or self.total is None:
            self.correct = hits.clone()
            self.total = torch.full((), logits.shape[0], dtype=torch.long, device=logits.device)
        else:
            self.correct.add_(hits)
            self.total.add_(logits.shape[0])

    def compute(self, name: str | None = None) -> Metrics:
        assert self.correct is not None and self.total is not None, "No updates were recorded"
        correct, total = torch.stack((self.correct.to(torch.float64), self.total.to(torch.float64))).tolist()  # one D2H
        return {self._key(name): correct / total}

    def clear(self) -> None:
        if self.correct is not None and self.total is not None:
            self.correct.zero_()
            self.total.zero_()

    def __call__(self, logits: torch.Tensor, target: torch.Tensor, threshold: float = 0.5) -> Scalar:
        preds = _hard_predictions(logits, threshold)
        return float((preds.reshape(-1) == _hard_targets(logits, target)).float().mean().item())


class _ConfusionMatrixMetric(Metric):
    """Base for metrics that are functions of the class confusion matrix (rows = truth, cols = prediction)."""

    def __init__(self, name: str) -> None:
        super().__init__(name)
        self.confusion: torch.Tensor | None = None

    def update(self, input: torch.Tensor, target: torch.Tensor) -> None:
        logits = input.detach()
        n_classes = 2 if logits.dim() == 1 or logits.shape[1] == 1 else logits.shape[1]
        preds = _hard_predictions(logits).reshape(-1)
        truth = _hard_targets(logits, target)
        flat = torch.bincount(truth * n_classes + preds, minlength=n_classes * n_classes)
        flat = flat.reshape(n_classes, n_classes)
        if self.confusion is None:
            self.confusion = flat.clone()
        else:
            self.confusion.add_(flat)

    def clear(self) -> None:
        if self.confusion is not None:
            self.confusion.zero_()

    def _matrix(self) -> np.ndarray:
        assert self.confusion is not None, "No updates were recorded"
        return self.confusion.cpu().numpy().astype(np.float64)

    def __call__(self, logits: torch.Tensor, target: torch.Tensor) -> Scalar:
        fresh = type(self).__new__(type(self))
        fresh.__dict__.update(self.__dict__)
        fresh.confusion = None
        fresh.update(logits, target)
        return next(iter(fresh.compute().values()))


class BalancedAccuracy(_ConfusionMatrixMetric):
    def __init__(self, name: str = "balanced_accuracy") -> None:
        super().__init__(name)

    def compute(self, name: str | None = None) -> Metrics:
        cm = self._matrix()
        support = cm.sum(axis=1)
        present = support > 0  # sklearn semantics: classes absent from y_true are ignored
        recall = np.diag(cm)[present] / support[present]
        return {self._key(name): float(recall.mean())}


class F1(_ConfusionMatrixMetric):
    def __init__(self, name: str = "F1 score", average: str | None = "weighted") -> None:
        super().__init__(name)
        self.average = average

    def compute(self, name: str | None = None) -> Metrics:
        cm = self._matrix()
        tp = np.diag(cm)
        support = cm.sum(axis=1)
        predicted = cm.sum(axis=0)
        if self.average == "micro":
            return {self._key(name): float(tp.sum() / max(cm.sum(), 1.0))}
        denom = support + predicted
        per_class = np.divide(2.0 * tp, denom, out=np.zeros_like(tp), where=denom > 0)
        seen = (support + predicted) > 0
        if self.average == "macro":
            value: Any = float(per_class[seen].mean())
        elif self.average == "weighted":
            value = float((per_class * support).sum() / max(support.sum(), 1.0))
        elif self.average is None:
            value = per_class[seen].tolist()
        elif self.average == "binary":
            value = float(per_class[1]) if len(per_class) > 1 else float(per_class[0])
        else:
            raise ValueError(f"Unsupported average: {self.average}")
        return {self._key(name): value}


class RocAuc(SimpleMetric):
    """Weighted one-vs-rest ROC-AUC from softmax probabilities (Mann-Whitney U, on device)."""

    def __init__(self, name: str = "ROC_AUC score") -> None:
        super().__init__(name)

    @staticmethod
    def _binary_auc(scores: torch.Tensor, positives: torch.Tensor) -> float:
        n_pos = int(positives.sum().item())
        n_neg = positives.numel() - n_pos
        if n_pos == 0 or n_neg == 0:
            raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
        # average ranks handle ties exactly as the trapezoidal ROC does
        order = torch.argsort(scores)
        sorted_scores = scores[order]
        ranks = torch.arange(1, scores.numel() + 1, dtype=torch.float64, device=scores.device)
        _, inverse, counts = torch.unique_consecutive(sorted_scores, return_inverse=True, return_counts=True)
        sums = torch.zeros(counts.numel(), dtype=torch.float64, device=scores.device).index_add_(0, inverse, ranks)
        avg_ranks = (sums / counts)[inverse]
        pos_rank_sum = avg_ranks[positives[order]].sum().item()
        return float((pos_rank_sum - n_pos * (n_pos + 1) / 2.0) / (n_pos * n_neg))

    def __call__(self, logits: torch.Tensor, target: torch.Tensor) -> Scalar:
        assert logits.shape[0] == target.shape[0], "Batch size of preds and targets do not match"
        prob = torch.softmax(logits.double(), dim=1)
        truth = _hard_targets(logits, target)
        if prob.shape[1] == 2:
            return self._binary_auc(prob[:, 1], truth == 1)
        total, weighted = 0, 0.0
        for cls in range(prob.shape[1]):
            mask = truth == cls
            support = int(mask.sum().item())
            if support == 0:
                continue
            weighted += support * self._binary_auc(prob[:, cls], mask)
            total += support
        return weighted / total


class BinarySoftDiceCoefficient(SimpleMetric):
    def __init__(
        self,
        name: str = "BinarySoftDiceCoefficient",
        epsilon: float = 1.0e-7,
        spatial_dimensions: tuple[int, ...] = (2, 3, 4),
        logits_threshold: float | None = 0.5,
    ) -> None:
        self.epsilon = epsilon
        self.spatial_dimensions = spatial_dimensions
        self.logits_threshold = logits_threshold
        super().__init__(name)

    def __call__(self, logits: torch.Tensor, target: torch.Tensor) -> Scalar:
        assert logits.shape[0] == target.shape[0], "Batch size of logits and targets do not match"
        assert logits.shape == target.shape, "Shapes of logits and target do not match"
        y_pred = (logits > self.logits_threshold).int() if self.logits_threshold else logits
        intersection = (y_pred * target).sum(dim=self.spatial_dimensions)
        union = 0.5 * (y_pred.sum(dim=self.spatial_dimensions) + target.sum(dim=self.spatial_dimensions))
        dice = intersection / (union + self.epsilon)
        return torch.mean(dice).item()


def compute_all(metrics: Sequence[Metric], name: str | None = None) -> Metrics:
    out: Metrics = {}
    for metric in metrics:
        out.update(metric.compute(name))
    return out
