"""Per-round record of losses and metrics."""

from __future__ import annotations

from functools import reduce
from pprint import pformat

from ..common.typing import Scalar


class History:
    def __init__(self) -> None:
        self.losses_distributed: list[tuple[int, float]] = []
        self.losses_centralized: list[tuple[int, float]] = []
        self.metrics_distributed_fit: dict[str, list[tuple[int, Scalar]]] = {}
        self.metrics_distributed: dict[str, list[tuple[int, Scalar]]] = {}
        self.metrics_centralized: dict[str, list[tuple[int, Scalar]]] = {}

    def add_loss_distributed(self, server_round: int, loss: float) -> None:
        self.losses_distributed.append((server_round, loss))

    def add_loss_centralized(self, server_round: int, loss: float) -> None:
        self.losses_centralized.append((server_round, loss))

    @staticmethod
    def _extend(store: dict[str, list[tuple[int, Scalar]]], server_round: int, metrics: dict[str, Scalar]) -> None:
        for key, value in metrics.items():
            store.setdefault(key, []).append((server_round, value))

    def add_metrics_distributed_fit(self, server_round: int, metrics: dict[str, Scalar]) -> None:
        self._extend(self.metrics_distributed_fit, server_round, metrics)

    def add_metrics_distributed(self, server_round: int, metrics: dict[str, Scalar]) -> None:
        self._extend(self.metrics_distributed, server_round, metrics)

    def add_metrics_centralized(self, server_round: int, metrics: dict[str, Scalar]) -> None:
        self._extend(self.metrics_centralized, server_round, metrics)

    def __repr__(self) -> str:
        parts = []
        if self.losses_distributed:
            parts.append("History (loss, distributed):\n" + reduce(
                lambda a, b: a + b, [f"\tround {r}: {loss}\n" for r, loss in self.losses_distributed]))
        if self.losses_centralized:
            parts.append("History (loss, centralized):\n" + reduce(
                lambda a, b: a + b, [f"\tround {r}: {loss}\n" for r, loss in self.losses_centralized]))
        for title, store in (("metrics, distributed, fit", self.metrics_distributed_fit),
                             ("metrics, distributed, evaluate", self.metrics_distributed),
                             ("metrics, centralized", self.metrics_centralized)):
            if store:
                parts.append(f"History ({title}):\n" + pformat(store))
        return "".join(parts)
