"""Helpers shared by exchangers: zero-copy state access and in-place state injection.

The reference converts every tensor to a CPU NumPy array on push and rebuilds an ``OrderedDict`` +
``load_state_dict(strict=True)`` on pull.  Here arrays stay on the device: push returns detached *views* (arena views
when the model lives in a ``ParameterArena``), pull copies in place (one flat copy when both sides are arena-shaped).
"""

from __future__ import annotations

from collections.abc import Iterable

import numpy as np
import torch
from torch import nn

from fl4health_b200.common.typing import NDArrays
from fl4health_b200.parallel.arena import arena_of


def state_views(model: nn.Module, names: Iterable[str] | None = None) -> NDArrays:
    arena = arena_of(model)
    if arena is not None:
        return arena.ndarrays(names)
    state = model.state_dict()
    keys = list(names) if names is not None else list(state.keys())
    return NDArrays([state[k].detach() for k in keys])


def inject_state(model: nn.Module, names: list[str], arrays: list, full: bool = False) -> None:
    """Copy ``arrays`` into the model's state entries ``names`` in place (shape-checked, dtype-cast)."""
    arena = arena_of(model)
    if arena is not None:
        arena.load_ndarrays(arrays, None if full else names)
        return
    state = model.state_dict(keep_vars=True)
    if full:
        assert len(arrays) == len(state), f"expected {len(state)} arrays for a full exchange, got {len(arrays)}"
    with torch.no_grad():
        for name, arr in zip(names, arrays):
            if name not in state:
                raise KeyError(f"Unexpected key {name} in exchanged parameters")
            dst = state[name]
            src = arr if isinstance(arr, torch.Tensor) else torch.from_numpy(np.asarray(arr, order="C").copy())
            if tuple(src.shape) != tuple(dst.shape):
                raise RuntimeError(f"size mismatch for {name}: {tuple(src.shape)} vs {tuple(dst.shape)}")
            dst.data.copy_(src.to(device=dst.device, dtype=dst.dtype), non_blocking=True)
