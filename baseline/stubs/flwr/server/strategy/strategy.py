from __future__ import annotations

from abc import ABC, abstractmethod

from ...common.typing import EvaluateIns, EvaluateRes, FitIns, FitRes, Parameters, Scalar
from ..client_manager import ClientManager
from ..client_proxy import ClientProxy


class Strategy(ABC):
    """What a server asks of an aggregation strategy, once per phase per round."""

    @abstractmethod
    def initialize_parameters(self, client_manager: ClientManager) -> Parameters | None: ...

    @abstractmethod
    def configure_fit(self, server_round: int, parameters: Parameters, client_manager: ClientManager) -> list[tuple[ClientProxy, FitIns]]: ...

    @abstractmethod
    def aggregate_fit(self, server_round: int, results: list[tuple[ClientProxy, FitRes]],
                      failures: list[tuple[ClientProxy, FitRes] | BaseException]) -> tuple[Parameters | None, dict[str, Scalar]]: ...

    @abstractmethod
    def configure_evaluate(self, server_round: int, parameters: Parameters, client_manager: ClientManager) -> list[tuple[ClientProxy, EvaluateIns]]: ...

    @abstractmethod
    def aggregate_evaluate(self, server_round: int, results: list[tuple[ClientProxy, EvaluateRes]],
                           failures: list[tuple[ClientProxy, EvaluateRes] | BaseException]) -> tuple[float | None, dict[str, Scalar]]: ...

    @abstractmethod
    def evaluate(self, server_round: int, parameters: Parameters) -> tuple[float, dict[str, Scalar]] | None: ...
