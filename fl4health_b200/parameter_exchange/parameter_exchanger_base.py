"""Exchanger interface (parity: ``fl4health/parameter_exchange/parameter_exchanger_base.py:8-20``)."""

from __future__ import annotations

from abc import ABC, abstractmethod

from torch import nn

from typing import TypeVar

from fl4health_b200.common.typing import Config, NDArrays


class ParameterExchanger(ABC):
    @abstractmethod
    def push_parameters(
        self, model: nn.Module, initial_model: nn.Module | None = None, config: Config | None = None
    ) -> NDArrays:
        """Model -> list of arrays sent to the server."""
        raise NotImplementedError

    @abstractmethod
    def pull_parameters(self, parameters: NDArrays, model: nn.Module, config: Config | None = None) -> None:
        """Server arrays -> model."""
        raise NotImplementedError


ExchangerType = TypeVar("ExchangerType", bound=ParameterExchanger)  # for code generic over the exchanger a client uses
