"""Seeding and RNG-state helpers (parity: ``fl4health/utils/random.py:11-116``).

The three host-side generators an FL run draws from (``random``, NumPy's legacy global generator, ``torch``) are kept in
one table, so seeding, un-seeding, snapshotting and restoring are each a loop over it."""

from __future__ import annotations

import random
import uuid
from collections.abc import Callable
from logging import INFO
from typing import Any, NamedTuple

import numpy as np
import torch

from fl4health_b200.common.logger import log


class _Source(NamedTuple):
    seed: Callable[[int], Any]
    unseed: Callable[[], Any]
    get_state: Callable[[], Any]
    set_state: Callable[[Any], Any]


def _seed_torch(seed: int) -> None:
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


_SOURCES: tuple[_Source, ...] = (
    _Source(random.seed, lambda: random.seed(None), random.getstate, random.setstate),
    _Source(np.random.seed, lambda: np.random.seed(None), lambda: np.random.get_state(legacy=False), np.random.set_state),
    _Source(_seed_torch, torch.seed, torch.get_rng_state, torch.set_rng_state),
)


def _seed_client_sampling(seed: int | None) -> None:
    """Client sampling draws from its own streams (``servers/client_manager.py``), which follow the global seed."""
    from fl4health_b200.servers.client_manager import sampling_streams

    sampling_streams.seed(seed)


def set_all_random_seeds(
    seed: int | None = 42, use_deterministic_torch_algos: bool = False, disable_torch_benchmarking: bool = False
) -> None:
    _seed_client_sampling(seed)
    if seed is not None:
        log(INFO, f"Setting seed to {seed}")
        for source in _SOURCES:
            source.seed(seed)
    else:
        log(INFO, "No seed provided. Using random seed.")
    if use_deterministic_torch_algos:
        log(INFO, "Setting torch.use_deterministic_algorithms to True.")
        torch.use_deterministic_algorithms(True)
    if disable_torch_benchmarking:
        log(INFO, "Disabling CUDNN benchmarking.")
        torch.backends.cudnn.benchmark = False


def unset_all_random_seeds() -> None:
    log(INFO, "Setting all random seeds to None. Reverting torch determinism settings")
    for source in _SOURCES:
        source.unseed()
    _seed_client_sampling(None)
    torch.use_deterministic_algorithms(False)


def save_random_state() -> tuple[tuple[Any, ...], dict[str, Any], torch.Tensor]:
    """(``random`` state, NumPy state, torch CPU generator state) — the argument order of ``restore_random_state``."""
    return tuple(source.get_state() for source in _SOURCES)  # type: ignore[return-value]


def restore_random_state(random_state: tuple[Any, ...], numpy_state: dict[str, Any], torch_state: torch.Tensor) -> None:
    for source, state in zip(_SOURCES, (random_state, numpy_state, torch_state)):
        source.set_state(state)


def generate_hash(length: int = 8) -> str:
    return uuid.uuid4().hex[:length]
