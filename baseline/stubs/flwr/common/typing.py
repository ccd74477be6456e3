"""Wire vocabulary (dataclasses) used between server, strategy and clients."""

from __future__ import annotations

from collections.abc import Callable
from dataclasses import dataclass, field
from enum import Enum
from typing import Any, Union

import numpy as np
import numpy.typing as npt

NDArray = npt.NDArray[Any]
NDArrays = list[NDArray]
Scalar = Union[bool, bytes, float, int, str]
Value = Union[bool, bytes, float, int, str, list[bool], list[bytes], list[float], list[int], list[str]]
Metrics = dict[str, Scalar]
MetricsAggregationFn = Callable[[list[tuple[int, Metrics]]], Metrics]
Config = dict[str, Scalar]
Properties = dict[str, Scalar]
UserConfig = dict[str, Scalar]


class Code(Enum):
    OK = 0
    GET_PROPERTIES_NOT_IMPLEMENTED = 1
    GET_PARAMETERS_NOT_IMPLEMENTED = 2
    FIT_NOT_IMPLEMENTED = 3
    EVALUATE_NOT_IMPLEMENTED = 4


@dataclass
class Status:
    code: Code
    message: str


@dataclass
class Parameters:
    tensors: list[bytes]
    tensor_type: str


@dataclass
class GetParametersIns:
    config: Config


@dataclass
class GetParametersRes:
    status: Status
    parameters: Parameters


@dataclass
class FitIns:
    parameters: Parameters
    config: dict[str, Scalar]


@dataclass
class FitRes:
    status: Status
    parameters: Parameters
    num_examples: int
    metrics: dict[str, Scalar] = field(default_factory=dict)


@dataclass
class EvaluateIns:
    parameters: Parameters
    config: dict[str, Scalar]


@dataclass
class EvaluateRes:
    status: Status
    loss: float
    num_examples: int
    metrics: dict[str, Scalar] = field(default_factory=dict)


@dataclass
class GetPropertiesIns:
    config: Config


@dataclass
class GetPropertiesRes:
    status: Status
    properties: Properties


@dataclass
class ReconnectIns:
    seconds: int | None


@dataclass
class DisconnectRes:
    reason: str


__all__ = [name for name in dir() if not name.startswith("_") and name not in ("np", "npt", "Any", "Union", "Enum")]
