"""Wire layouts for side information appended to a weight list.

Parity: ``fl4health/parameter_exchange/parameter_packer.py:13-142`` — the list layouts are preserved exactly
(``weights ++ variates``; ``weights ++ [scalar]``; ``weights ++ [names]``; ``values ++ indices ++ shapes ++ [names]``)
because strategies and the server checkpoint modules split on them.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Generic, TypeVar

import numpy as np
import torch

from fl4health_b200.common.typing import NDArray, NDArrays

T = TypeVar("T")


def _slice(arrays: list, start: int | None, stop: int | None) -> NDArrays:
    if isinstance(arrays, NDArrays):
        return arrays.sliced(start, stop)
    return NDArrays(arrays[slice(start, stop)])


def _scalar(value: object) -> float:
    if isinstance(value, torch.Tensor):
        return float(value.item())
    return float(np.asarray(value).item())


class ParameterPacker(ABC, Generic[T]):
    @abstractmethod
    def pack_parameters(self, model_weights: NDArrays, additional_parameters: T) -> NDArrays:
        raise NotImplementedError

    @abstractmethod
    def unpack_parameters(self, packed_parameters: NDArrays) -> tuple[NDArrays, T]:
        raise NotImplementedError


class ParameterPackerWithControlVariates(ParameterPacker[NDArrays]):
    def __init__(self, size_of_model_params: int) -> None:
        self.size_of_model_params = size_of_model_params

    def pack_parameters(self, model_weights: NDArrays, additional_parameters: NDArrays) -> NDArrays:
        packed = NDArrays(list(model_weights) + list(additional_parameters))
        packed.flat, packed.layout = getattr(model_weights, "flat", None), getattr(model_weights, "layout", None)
        packed.int_flat = getattr(model_weights, "int_flat", None)
        packed.aux_flat = getattr(additional_parameters, "flat", None)
        packed.aux_layout = getattr(additional_parameters, "layout", None)
        return packed

    def unpack_parameters(self, packed_parameters: NDArrays) -> tuple[NDArrays, NDArrays]:
        split = self.size_of_model_params
        return _slice(packed_parameters, None, split), _slice(packed_parameters, split, None)


class _TrailingScalarPacker(ParameterPacker[float]):
    def pack_parameters(self, model_weights: NDArrays, additional_parameters: float) -> NDArrays:
        # a device-resident scalar (e.g. the clipping bit written by the clip kernel) travels as is: no host read-back
        trailing = additional_parameters if isinstance(additional_parameters, torch.Tensor) else np.array(additional_parameters)
        packed = NDArrays(list(model_weights) + [trailing])
        # keep arena metadata of the weight part so fused aggregation still applies
        packed.flat, packed.layout = getattr(model_weights, "flat", None), getattr(model_weights, "layout", None)
        packed.int_flat = getattr(model_weights, "int_flat", None)
        return packed

    def unpack_parameters(self, packed_parameters: NDArrays) -> tuple[NDArrays, float]:
        assert len(packed_parameters) >= 1
        return _slice(packed_parameters, None, -1), _scalar(packed_parameters[-1])


class ParameterPackerWithClippingBit(_TrailingScalarPacker):
    """``weights ++ [clipping_bit_or_bound]``."""


class ParameterPackerAdaptiveConstraint(_TrailingScalarPacker):
    """``weights ++ [mu]`` (server -> client) or ``weights ++ [train_loss]`` (client -> server)."""


class ParameterPackerWithLayerNames(ParameterPacker[list[str]]):
    def pack_parameters(self, model_weights: NDArrays, weights_names: list[str]) -> NDArrays:
        return NDArrays(list(model_weights) + [np.array(weights_names)])

    def unpack_parameters(self, packed_parameters: NDArrays) -> tuple[NDArrays, list[str]]:
        names = packed_parameters[-1]
        names_list = names.tolist() if isinstance(names, np.ndarray) else list(names)
        return _slice(packed_parameters, None, -1), [str(n) for n in names_list]


class SparseCooParameterPacker(ParameterPacker[tuple[NDArrays, NDArrays, list[str]]]):
    def pack_parameters(
        self, model_parameters: NDArrays, additional_parameters: tuple[NDArrays, NDArrays, list[str]]
    ) -> NDArrays:
        parameter_indices, tensor_shapes, tensor_names = additional_parameters
        return NDArrays(
            list(model_parameters) + list(parameter_indices) + list(tensor_shapes) + [np.array(tensor_names)]
        )

    def unpack_parameters(
        self, packed_parameters: NDArrays
    ) -> tuple[NDArrays, tuple[NDArrays, NDArrays, list[str]]]:
        assert len(packed_parameters) % 3 == 1
        split = (len(packed_parameters) - 1) // 3
        values = NDArrays(packed_parameters[:split])
        indices = NDArrays(packed_parameters[split : 2 * split])
        shapes = NDArrays(packed_parameters[2 * split : 3 * split])
        names = packed_parameters[3 * split]
        names_list = names.tolist() if isinstance(names, np.ndarray) else list(names)
        return values, (indices, shapes, [str(n) for n in names_list])

    @staticmethod
    def extract_coo_info_from_dense(x: torch.Tensor) -> tuple[NDArray, NDArray, NDArray]:
        """(values, [nnz, ndim] indices, shape) of the non-zero entries of ``x`` — stays on ``x``'s device."""
        indices = torch.nonzero(x, as_tuple=False)
        values = x[tuple(indices.t())] if indices.numel() > 0 else x.new_zeros((0,))
        return values, indices, np.array(list(x.shape))
