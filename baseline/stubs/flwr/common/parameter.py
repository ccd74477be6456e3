"""ndarray <-> bytes conversion (``np.save`` blobs, no pickling), as Flower's default serialiser does."""

from __future__ import annotations

from io import BytesIO
from typing import cast

import numpy as np

from .typing import NDArray, NDArrays, Parameters


def ndarray_to_bytes(ndarray: NDArray) -> bytes:
    buffer = BytesIO()
    np.save(buffer, ndarray, allow_pickle=False)
    return buffer.getvalue()


def bytes_to_ndarray(tensor: bytes) -> NDArray:
    return cast(NDArray, np.load(BytesIO(tensor), allow_pickle=False))


def ndarrays_to_parameters(ndarrays: NDArrays) -> Parameters:
    return Parameters(tensors=[ndarray_to_bytes(a) for a in ndarrays], tensor_type="numpy.ndarray")


def parameters_to_ndarrays(parameters: Parameters) -> NDArrays:
    return [bytes_to_ndarray(t) for t in parameters.tensors]
