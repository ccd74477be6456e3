"""Full-model exchange (parity: ``fl4health/parameter_exchange/full_exchanger.py:10-47``).

All ``state_dict()`` entries in key order, including BatchNorm running statistics and integer counters.
"""

from __future__ import annotations

from torch import nn

from fl4health_b200.common.typing import Config, NDArrays
from fl4health_b200.parallel.arena import arena_of
from fl4health_b200.parameter_exchange._state import inject_state, state_views
from fl4health_b200.parameter_exchange.parameter_exchanger_base import ParameterExchanger


class FullParameterExchanger(ParameterExchanger):
    def push_parameters(
        self, model: nn.Module, initial_model: nn.Module | None = None, config: Config | None = None
    ) -> NDArrays:
        return state_views(model)

    def pull_parameters(self, parameters: NDArrays, model: nn.Module, config: Config | None = None) -> None:
        arena = arena_of(model)
        if arena is not None:  # keys are cached in the arena: no state_dict() walk on the per-round path
            arena.load_ndarrays(parameters)
            return
        inject_state(model, list(model.state_dict().keys()), parameters, full=True)
