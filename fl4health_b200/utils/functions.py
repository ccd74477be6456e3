"""Small numerical helpers (parity: ``fl4health/utils/functions.py:10-108``)."""

from __future__ import annotations

from typing import Any

import numpy as np
import torch

from fl4health_b200.common.typing import FitRes, NDArrays, parameters_to_ndarrays, to_numpy


class BernoulliSample(torch.autograd.Function):
    """Straight-through Bernoulli sampling: forward ``bernoulli(p)``, backward ``p * grad`` (FedPM masked layers)."""

    @staticmethod
    def forward(ctx: Any, bernoulli_probs: torch.Tensor) -> torch.Tensor:  # type: ignore[override]
        ctx.save_for_backward(bernoulli_probs)
        return torch.bernoulli(bernoulli_probs)

    @staticmethod
    def backward(ctx: Any, grad_output: torch.Tensor) -> torch.Tensor:  # type: ignore[override]
        (bernoulli_probs,) = ctx.saved_tensors
        return bernoulli_probs * grad_output


def bernoulli_sample(bernoulli_probs: torch.Tensor) -> torch.Tensor:
    return BernoulliSample.apply(bernoulli_probs)  # type: ignore[no-any-return]


def sigmoid_inverse(x: torch.Tensor) -> torch.Tensor:
    return -torch.log(1.0 / x - 1.0)


def select_zeroeth_element(array: Any) -> float:
    """First element of an arbitrarily shaped array (used by the pseudo sort)."""
    arr = array if isinstance(array, np.ndarray) else to_numpy(array)
    return float(arr.reshape(-1)[0]) if arr.size > 0 else 0.0


def pseudo_sort_scoring_function(client_result: tuple[Any, NDArrays, int]) -> float:
    """Deterministic-ish score: sum of the first element of every array plus the sample count."""
    _, client_arrays, sample_count = client_result
    total = 0.0
    for arr in client_arrays:
        floating = arr.dtype.kind == "f" if isinstance(arr, np.ndarray) else getattr(arr, "is_floating_point", lambda: True)()
        if floating:  # layer names, integer counters and index arrays do not take part
            total += select_zeroeth_element(arr)
    return total + sample_count


def decode_and_pseudo_sort_results(
    results: list[tuple[Any, FitRes]], materialize: bool = True
) -> list[tuple[Any, NDArrays, int]]:
    """(proxy, arrays, n) triples in a canonical order.

    The reference sorts by a numeric pseudo-score because Flower client ids are random UUIDs
    (``fl4health/utils/functions.py:84-108``).  Here client ids are stable (rank / client name), so ordering by
    ``cid`` gives bit-deterministic fixed-order summation without touching (or syncing on) the payload.
    """
    ordered = sorted(results, key=lambda item: str(getattr(item[0], "cid", "")))
    decoded = []
    for proxy, res in ordered:
        arrays = parameters_to_ndarrays(res.parameters)
        if materialize and getattr(arrays, "ctx", None) is not None and arrays.ctx.world_size > 1:
            # SPMD fallback for strategies that need every client's full payload: broadcast from the owner.  The OWNER
            # has to take part too (its copy is the broadcast source) — testing `remote` alone left it out and
            # deadlocked every strategy on this path.
            from fl4health_b200.parallel.spmd import materialize as _materialize

            arrays = _materialize(arrays)
        decoded.append((proxy, arrays, res.num_examples))
    return decoded
