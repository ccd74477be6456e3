"""Metric interface (parity: ``fl4health/metrics/base_metrics.py:8-67``)."""

from __future__ import annotations

from abc import ABC, abstractmethod
from enum import Enum

import torch

from fl4health_b200.common.typing import Metrics

TEST_NUM_EXAMPLES_KEY = "test - num_examples"
TEST_LOSS_KEY = "test - checkpoint"


class MetricPrefix(Enum):
    TEST_PREFIX = "test -"
    VAL_PREFIX = "val -"


class Metric(ABC):
    def __init__(self, name: str) -> None:
        self.name = name

    @abstractmethod
    def update(self, input: torch.Tensor, target: torch.Tensor) -> None:
        raise NotImplementedError

    @abstractmethod
    def compute(self, name: str | None = None) -> Metrics:
        raise NotImplementedError

    @abstractmethod
    def clear(self) -> None:
        raise NotImplementedError

    def _key(self, name: str | None) -> str:
        return f"{name} - {self.name}" if name is not None else self.name
