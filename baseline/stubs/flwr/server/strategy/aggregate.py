"""Host-side (NumPy) aggregation helpers."""

from __future__ import annotations

from functools import reduce

import numpy as np

from ...common.typing import NDArrays


def aggregate(results: list[tuple[NDArrays, int]]) -> NDArrays:
    """Example-count weighted mean, layer by layer."""
    total = sum(n for _, n in results)
    scaled = [[layer * n for layer in weights] for weights, n in results]
    return [reduce(np.add, layers) / total for layers in zip(*scaled)]


def weighted_loss_avg(results: list[tuple[int, float]]) -> float:
    total = sum(n for n, _ in results)
    return sum(n * loss for n, loss in results) / total
