"""Wire types of the engine.

These play the role of ``flwr.common`` dataclasses but are *zero-copy*: a ``Parameters``
object carries references to live ``torch.Tensor``s (usually views into a rank's flat device arena) or numpy
arrays, never serialized bytes.  The np.save/protobuf/gRPC path of the reference
(``fl4health/parameter_exchange/full_exchanger.py:30``) simply does not exist here.
"""

from __future__ import annotations

from collections.abc import Callable
from dataclasses import dataclass, field
from enum import Enum
from typing import Any

import numpy as np
import torch

Scalar = bool | bytes | float | int | str
Config = dict[str, Scalar]
Metrics = dict[str, Scalar]
Properties = dict[str, Scalar]
NDArray = np.ndarray | torch.Tensor
MetricsAggregationFn = Callable[[list[tuple[int, Metrics]]], Metrics]


class NDArrays(list):
    """A list of arrays with optional arena metadata.

    When every entry is a view into one contiguous flat buffer, ``flat`` holds that buffer and ``layout`` the
    (offset, numel, shape) triples; strategies and exchangers use it to replace per-layer loops with one fused
    kernel over the flat storage.  It behaves as a plain ``list`` otherwise.
    """

    flat: torch.Tensor | None
    layout: Any
    # second flat region for packed side payloads that are themselves arena-shaped (SCAFFOLD variates)
    aux_flat: torch.Tensor | None = None
    aux_layout: Any = None
    # all integer entries (``num_batches_tracked`` ...) as one int64 tensor, in list order, when available
    int_flat: torch.Tensor | None = None
    # a NAMED SUBSET of an arena (partial exchange: FedPer, FedRep, FedBN ...): the arena-shaped buffer the entries are
    # views of, the arena (layout) and the state keys, so subsets can ride the whole-arena kernels
    subset_flat: torch.Tensor | None = None
    subset_layout: Any = None
    subset_names: tuple[str, ...] | None = None

    def __init__(self, iterable: Any = (), flat: torch.Tensor | None = None, layout: Any = None) -> None:
        super().__init__(iterable)
        self.flat = flat
        self.layout = layout

    def sliced(self, start: int | None, stop: int | None) -> NDArrays:
        """``self[start:stop]`` that keeps arena / ownership metadata (packers use it to split side information
        from weights without losing the fused-aggregation fast path)."""
        part = NDArrays(list.__getitem__(self, slice(start, stop)))
        layout = self.layout
        if self.flat is not None and layout is not None and len(part) == len(layout.state_keys) and (start in (None, 0)):
            part.flat, part.layout, part.int_flat = self.flat, layout, self.int_flat
        elif (
            self.aux_flat is not None and self.aux_layout is not None and layout is not None
            and start == len(layout.state_keys) and len(part) == len(self.aux_layout.state_keys)
        ):
            part.flat, part.layout = self.aux_flat, self.aux_layout
        return part


class Code(Enum):
    OK = 0
    GET_PROPERTIES_NOT_IMPLEMENTED = 1
    GET_PARAMETERS_NOT_IMPLEMENTED = 2
    FIT_NOT_IMPLEMENTED = 3
    EVALUATE_NOT_IMPLEMENTED = 4


@dataclass
class Status:
    code: Code = Code.OK
    message: str = "Success"


@dataclass
class Parameters:
    tensors: list[Any]
    tensor_type: str = "torch"
    # arena metadata (optional): one contiguous buffer backing every entry of ``tensors``.
    flat: Any = None
    layout: Any = None
    int_flat: Any = None  # all integer entries as one int64 tensor (see ``NDArrays.int_flat``)
    aux_flat: Any = None  # second arena-shaped block packed behind the model state (see ``NDArrays.aux_flat``)
    aux_layout: Any = None
    subset_flat: Any = None  # named arena subset (see ``NDArrays.subset_flat``)
    subset_layout: Any = None
    subset_names: Any = None


def ndarrays_to_parameters(ndarrays: list[NDArray] | NDArrays) -> Parameters:
    return Parameters(
        tensors=list(ndarrays),
        tensor_type="torch",
        flat=getattr(ndarrays, "flat", None),
        layout=getattr(ndarrays, "layout", None),
        int_flat=getattr(ndarrays, "int_flat", None),
        aux_flat=getattr(ndarrays, "aux_flat", None),
        aux_layout=getattr(ndarrays, "aux_layout", None),
        subset_flat=getattr(ndarrays, "subset_flat", None),
        subset_layout=getattr(ndarrays, "subset_layout", None),
        subset_names=getattr(ndarrays, "subset_names", None),
    )


def parameters_to_ndarrays(parameters: Parameters) -> NDArrays:
    tagged = getattr(parameters, "_arrays", None)
    if tagged is not None:  # SPMD transport: keep ownership tags / remote placeholders intact
        return tagged
    arrays = NDArrays(parameters.tensors, flat=parameters.flat, layout=parameters.layout)
    arrays.int_flat = getattr(parameters, "int_flat", None)
    arrays.aux_flat, arrays.aux_layout = getattr(parameters, "aux_flat", None), getattr(parameters, "aux_layout", None)
    for tag in ("subset_flat", "subset_layout", "subset_names"):
        setattr(arrays, tag, getattr(parameters, tag, None))
    return arrays


def to_numpy(array: NDArray) -> np.ndarray:
    if isinstance(array, torch.Tensor):
        return array.detach().cpu().numpy()
    return np.asarray(array)


def to_tensor(array: NDArray, device: torch.device | str | None = None) -> torch.Tensor:
    if isinstance(array, torch.Tensor):
        return array if device is None else array.to(device, non_blocking=True)
    arr = np.asarray(array)
    if arr.dtype.kind in ("U", "S", "O"):
        raise TypeError("string arrays cannot be converted to tensors")
    t = torch.from_numpy(np.asarray(arr, order="C").copy())
    return t if device is None else t.to(device, non_blocking=True)


@dataclass
class FitIns:
    parameters: Parameters
    config: Config = field(default_factory=dict)


@dataclass
class FitRes:
    status: Status
    parameters: Parameters
    num_examples: int
    metrics: Metrics = field(default_factory=dict)


@dataclass
class EvaluateIns:
    parameters: Parameters
    config: Config = field(default_factory=dict)


@dataclass
class EvaluateRes:
    status: Status
    loss: float
    num_examples: int
    metrics: Metrics = field(default_factory=dict)


@dataclass
class GetPropertiesIns:
    config: Config = field(default_factory=dict)


@dataclass
class GetPropertiesRes:
    status: Status
    properties: Properties = field(default_factory=dict)


@dataclass
class GetParametersIns:
    config: Config = field(default_factory=dict)


@dataclass
class GetParametersRes:
    status: Status
    parameters: Parameters


@dataclass
class ReconnectIns:
    seconds: int | None = None


@dataclass
class DisconnectRes:
    reason: str = ""
