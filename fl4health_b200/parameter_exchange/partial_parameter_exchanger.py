"""Base for exchangers that pick a per-round subset (parity: ``partial_parameter_exchanger.py:14-41``)."""

from __future__ import annotations

from abc import abstractmethod
from typing import Generic, TypeVar

from torch import nn

from fl4health_b200.common.typing import NDArrays
from fl4health_b200.parameter_exchange.parameter_exchanger_base import ParameterExchanger
from fl4health_b200.parameter_exchange.parameter_packer import ParameterPacker

T = TypeVar("T")


class PartialParameterExchanger(ParameterExchanger, Generic[T]):
    parameter_packer: ParameterPacker[T]

    def __init__(self, parameter_packer: ParameterPacker[T]) -> None:
        self.parameter_packer = parameter_packer

    def pack_parameters(self, model_weights: NDArrays, additional_parameters: T) -> NDArrays:
        return self.parameter_packer.pack_parameters(model_weights, additional_parameters)

    def unpack_parameters(self, packed_parameters: NDArrays) -> tuple[NDArrays, T]:
        return self.parameter_packer.unpack_parameters(packed_parameters)

    @abstractmethod
    def select_parameters(self, model: nn.Module, initial_model: nn.Module | None = None) -> tuple[NDArrays, T]:
        raise NotImplementedError
