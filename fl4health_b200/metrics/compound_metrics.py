"""Metric wrappers (parity: ``fl4health/metrics/compound_metrics.py:17-167``)."""

from __future__ import annotations

import copy
from collections.abc import Callable, Sequence
from logging import WARNING
from typing import Generic, TypeVar

import torch

from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Metrics
from fl4health_b200.metrics.base_metrics import Metric

T = TypeVar("T", bound=Metric)
TorchTransformFunction = Callable[[torch.Tensor], torch.Tensor]


class EmaMetric(Metric, Generic[T]):
    """Exponential moving average over successive ``compute`` results of the wrapped metric:
    ``s_t = a * m_t + (1 - a) * s_{t-1}``; the first score is stored as is.  ``clear`` resets the inner accumulation
    but not the running average."""

    def __init__(self, metric: T, smoothing_factor: float = 0.1, name: str | None = None) -> None:
        self.metric = copy.deepcopy(metric)
        assert 0.0 <= smoothing_factor <= 1.0, f"smoothing_factor should be in [0, 1] but was {smoothing_factor}"
        self.smoothing_factor = smoothing_factor
        self.previous_score: Metrics | None = None
        super().__init__(f"EMA_{self.metric.name}" if name is None else name)

    def update(self, input: torch.Tensor, target: torch.Tensor) -> None:
        self.metric.update(input, target)

    def compute(self, name: str | None = None) -> Metrics:
        inner_name = self.metric.name
        self.metric.name = self.name
        try:
            current = self.metric.compute(name)
        finally:
            self.metric.name = inner_name
        if self.previous_score is None:
            self.previous_score = {}
            for key, score in current.items():
                if isinstance(score, (int, float)) and not isinstance(score, bool):
                    self.previous_score[key] = score
                else:
                    log(WARNING, "EMAMetric is only compatible with float or int metrics, but metrics contains a value "
                                 f"with type: {type(score)} at key: {key}. These values will be ignored in subsequent computations.")
            return dict(self.previous_score)
        a = self.smoothing_factor
        for key, previous in self.previous_score.items():
            score = current[key]
            if not isinstance(score, (str, bytes)) and not isinstance(previous, (str, bytes)):
                self.previous_score[key] = a * score + (1 - a) * previous
        return dict(self.previous_score)

    def clear(self) -> None:
        self.metric.clear()


class TransformsMetric(Metric, Generic[T]):
    """Applies tensor transforms (in order) to predictions / targets before updating the wrapped metric."""

    def __init__(
        self, metric: T, pred_transforms: Sequence[TorchTransformFunction] | None = None,
        target_transforms: Sequence[TorchTransformFunction] | None = None,
    ) -> None:
        self.metric = copy.deepcopy(metric)
        self.pred_transforms = list(pred_transforms or [])
        self.target_transforms = list(target_transforms or [])
        super().__init__(name=self.metric.name)

    def update(self, pred: torch.Tensor, target: torch.Tensor) -> None:
        for transform in self.pred_transforms:
            pred = transform(pred)
        for transform in self.target_transforms:
            target = transform(target)
        self.metric.update(pred, target)

    def compute(self, name: str | None = None) -> Metrics:
        return self.metric.compute(name)

    def clear(self) -> None:
        self.metric.clear()
