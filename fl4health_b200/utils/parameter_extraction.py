"""Model -> wire ``Parameters`` helpers (parity: ``fl4health/utils/parameter_extraction.py``).  Tensors stay on their
device (arena views when available); no NumPy round trip."""

from __future__ import annotations

from collections.abc import Iterable

import torch
from torch import nn

from fl4health_b200.common.typing import Parameters, ndarrays_to_parameters
from fl4health_b200.parameter_exchange._state import state_views


def get_all_model_parameters(model: nn.Module) -> Parameters:
    return ndarrays_to_parameters(state_views(model))


def check_shape_match(params1: Iterable[torch.Tensor], params2: Iterable[torch.Tensor], error_message: str) -> None:
    first, second = list(params1), list(params2)
    assert len(first) == len(second), f"Parameter length mismatch: {len(first)} vs {len(second)}. {error_message}"
    for a, b in zip(first, second):
        assert a.shape == b.shape, error_message
