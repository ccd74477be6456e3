"""Per-split metric bookkeeping (parity: ``fl4health/metrics/metric_managers.py:11-86``).

A manager is handed prototype metrics once and a dictionary of predictions per batch.  On the first batch it *binds* a
private copy of every prototype to every prediction key; from then on an update is a walk over that flat binding list.
Result keys keep the reference's schema ``"{manager} - {prediction_key} - {metric}"``.  ``clear`` resets the bound
metric objects *in place* (device-side counters referenced by a captured CUDA graph keep their addresses); ``reset``
drops the bindings, so the next batch may bring different prediction keys.
"""

from __future__ import annotations

import copy
from collections.abc import Sequence

import torch

from fl4health_b200.common.typing import Metrics
from fl4health_b200.metrics.base_metrics import Metric
from fl4health_b200.utils.typing import TorchPredType, TorchTargetType

_KEY_MISMATCH = (
    "Received a dict with multiple targets, but the keys of the targets do not match the keys of the "
    "predictions. Please pass a single target or ensure the keys between preds and target are the same"
)


class MetricManager:
    def __init__(self, metrics: Sequence[Metric], metric_manager_name: str) -> None:
        self.original_metrics = metrics
        self.metric_manager_name = metric_manager_name
        self._bound: list[tuple[str, Metric]] = []  # (prediction key, this key's own copy of a prototype), update order
        self._bound_keys: tuple[str, ...] | None = None  # None: nothing bound yet (a manager may have no metrics at all)

    @property
    def metrics_per_prediction_type(self) -> dict[str, Sequence[Metric]]:
        """The bindings grouped by prediction key (the reference's attribute)."""
        grouped: dict[str, list[Metric]] = {key: [] for key in self._bound_keys or ()}
        for key, metric in self._bound:
            grouped.setdefault(key, []).append(metric)
        return grouped  # type: ignore[return-value]

    def _bind(self, prediction_keys: Sequence[str]) -> None:
        self._bound_keys = tuple(prediction_keys)
        self._bound = [(key, copy.deepcopy(prototype)) for key in prediction_keys for prototype in self.original_metrics]

    @staticmethod
    def _target_for(key: str, target: TorchTargetType) -> torch.Tensor:
        if isinstance(target, torch.Tensor):
            return target
        return next(iter(target.values())) if len(target) == 1 else target[key]

    def update(self, preds: TorchPredType, target: TorchTargetType) -> None:
        if self._bound_keys is None:
            self._bind(list(preds))
        if isinstance(target, dict) and len(target) > 1:
            self.check_target_prediction_keys_equal(preds, target)
        assert set(self._bound_keys or ()) == set(preds), "prediction keys changed between batches: call reset() first"
        for key, metric in self._bound:
            metric.update(preds[key], self._target_for(key, target))

    def compute(self) -> Metrics:
        results: Metrics = {}
        for key, metric in self._bound:
            results.update(metric.compute(f"{self.metric_manager_name} - {key}"))
        return results

    def clear(self) -> None:
        for _, metric in self._bound:
            metric.clear()

    def reset(self) -> None:
        self._bound, self._bound_keys = [], None

    def check_target_prediction_keys_equal(self, preds: dict[str, torch.Tensor], target: dict[str, torch.Tensor]) -> None:
        assert target.keys() == preds.keys(), _KEY_MISMATCH
