"""Per-split metric bookkeeping (parity: ``fl4health/metrics/metric_managers.py:11-86``).

Result keys keep the reference's schema ``"{manager} - {prediction_key} - {metric}"``.  Unlike the reference,
``clear`` resets the metric objects *in place* when possible (``Metric.clear``) so device-side counters referenced
by a captured CUDA graph keep their addresses; ``reset`` drops them entirely.
"""

from __future__ import annotations

import copy
from collections.abc import Sequence

import torch

from fl4health_b200.common.typing import Metrics
from fl4health_b200.metrics.base_metrics import Metric
from fl4health_b200.utils.typing import TorchPredType, TorchTargetType


class MetricManager:
    def __init__(self, metrics: Sequence[Metric], metric_manager_name: str) -> None:
        self.original_metrics = metrics
        self.metric_manager_name = metric_manager_name
        self.metrics_per_prediction_type: dict[str, Sequence[Metric]] = {}

    def update(self, preds: TorchPredType, target: TorchTargetType) -> None:
        if not self.metrics_per_prediction_type:
            self.metrics_per_prediction_type = {key: copy.deepcopy(self.original_metrics) for key in preds}
        if isinstance(target, dict):
            if len(target) > 1:
                self.check_target_prediction_keys_equal(preds, target)
            else:
                target = next(iter(target.values()))
        assert len(preds) == len(self.metrics_per_prediction_type)
        for prediction_key, pred in preds.items():
            tgt = target if isinstance(target, torch.Tensor) else target[prediction_key]
            for metric in self.metrics_per_prediction_type[prediction_key]:
                metric.update(pred, tgt)

    def compute(self) -> Metrics:
        results: Metrics = {}
        for prediction_key, metrics in self.metrics_per_prediction_type.items():
            for metric in metrics:
                results.update(metric.compute(f"{self.metric_manager_name} - {prediction_key}"))
        return results

    def clear(self) -> None:
        for metrics in self.metrics_per_prediction_type.values():
            for metric in metrics:
                metric.clear()

    def reset(self) -> None:
        self.metrics_per_prediction_type = {}

    def check_target_prediction_keys_equal(
        self, preds: dict[str, torch.Tensor], target: dict[str, torch.Tensor]
    ) -> None:
        assert target.keys() == preds.keys(), (
            "Received a dict with multiple targets, but the keys of the targets do not match the keys of the "
            "predictions. Please pass a single target or ensure the keys between preds and target are the same"
        )
