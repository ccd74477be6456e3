"""Seeding and RNG-state helpers (parity: ``fl4health/utils/random.py:11-116``)."""

from __future__ import annotations

import random
import uuid
from logging import INFO
from typing import Any

import numpy as np
import torch

from fl4health_b200.common.logger import log


def set_all_random_seeds(
    seed: int | None = 42, use_deterministic_torch_algos: bool = False, disable_torch_benchmarking: bool = False
) -> None:
    if seed is None:
        log(INFO, "No seed provided. Using random seed.")
    else:
        log(INFO, f"Setting seed to {seed}")
        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)
        if torch.cuda.is_available():
            torch.cuda.manual_seed_all(seed)
    if use_deterministic_torch_algos:
        log(INFO, "Setting torch.use_deterministic_algorithms to True.")
        torch.use_deterministic_algorithms(True)
    if disable_torch_benchmarking:
        log(INFO, "Disabling CUDNN benchmarking.")
        torch.backends.cudnn.benchmark = False


def unset_all_random_seeds() -> None:
    log(INFO, "Setting all random seeds to None. Reverting torch determinism settings")
    random.seed(None)
    np.random.seed(None)
    torch.seed()
    torch.use_deterministic_algorithms(False)


def save_random_state() -> tuple[tuple[Any, ...], dict[str, Any], torch.Tensor]:
    return random.getstate(), np.random.get_state(legacy=False), torch.get_rng_state()  # type: ignore[return-value]


def restore_random_state(
    random_state: tuple[Any, ...], numpy_state: dict[str, Any], torch_state: torch.Tensor
) -> None:
    random.setstate(random_state)  # type: ignore[arg-type]
    np.random.set_state(numpy_state)  # type: ignore[arg-type]
    torch.set_rng_state(torch_state)


def generate_hash(length: int = 8) -> str:
    return str(uuid.uuid4()).replace("-", "")[:length]
