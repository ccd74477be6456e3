"""Shared type aliases (parity: ``fl4health/utils/typing.py:1-33``)."""

from __future__ import annotations

import logging
from collections.abc import Callable
from enum import Enum
from typing import TYPE_CHECKING

import torch
from torch import nn

from fl4health_b200.common.typing import EvaluateRes, FitRes, NDArrays

if TYPE_CHECKING:
    from fl4health_b200.servers.client_proxy import ClientProxy

TorchInputType = torch.Tensor | dict[str, torch.Tensor]
TorchTargetType = torch.Tensor | dict[str, torch.Tensor]
TorchPredType = dict[str, torch.Tensor]
TorchFeatureType = dict[str, torch.Tensor]
TorchTransformFunction = Callable[[torch.Tensor], torch.Tensor]
LayerSelectionFunction = Callable[[nn.Module, nn.Module | None], tuple[NDArrays, list[str]]]

FitFailures = list["tuple[ClientProxy, FitRes] | BaseException"]
EvaluateFailures = list["tuple[ClientProxy, EvaluateRes] | BaseException"]


class LogLevel(Enum):
    NOTSET = logging.NOTSET
    DEBUG = logging.DEBUG
    INFO = logging.INFO
    WARNING = logging.WARNING
    ERROR = logging.ERROR
    CRITICAL = logging.CRITICAL
