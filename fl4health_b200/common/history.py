"""Training history container (role of ``flwr.server.history.History``; pickled inside server state)."""

from __future__ import annotations

from functools import reduce
from pprint import pformat

from fl4health_b200.common.typing import Scalar


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
        rep = ""
        if self.losses_distributed:
            rep += "History (loss, distributed):\n" + reduce(
                lambda a, b: a + b, [f"\tround {r}: {loss}\n" for r, loss in self.losses_distributed]
            )
        if self.losses_centralized:
            rep += "History (loss, centralized):\n" + reduce(
                lambda a, b: a + b, [f"\tround {r}: {loss}\n" for r, loss in self.losses_centralized]
            )
        if self.metrics_distributed_fit:
            rep += "History (metrics, distributed, fit):\n" + pformat(self.metrics_distributed_fit) + "\n"
        if self.metrics_distributed:
            rep += "History (metrics, distributed, evaluate):\n" + pformat(self.metrics_distributed) + "\n"
        if self.metrics_centralized:
            rep += "History (metrics, centralized):\n" + pformat(self.metrics_centralized) + "\n"
        return rep
