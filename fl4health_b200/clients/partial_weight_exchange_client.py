"""Client for partial (per-round selected) weight exchange (parity: ``partial_weight_exchange_client.py:18-148``): keeps
an ``initial_model`` copy of the round-start weights so selection rules can score drift."""

from __future__ import annotations

import copy
from typing import Any

from torch import nn

from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.common.typing import Config, NDArrays
from fl4health_b200.parameter_exchange.parameter_exchanger_base import ParameterExchanger


class PartialWeightExchangeClient(BasicClient):
    """``store_initial_model=True`` keeps a frozen twin holding the weights as they were right after the last pull; it is
    handed to the exchanger's ``push_parameters`` (drift-based selection criteria compare against it)."""

    def __init__(self, *args: Any, store_initial_model: bool = False, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self.store_initial_model = store_initial_model
        self.initial_model: nn.Module | None = None

    def get_parameter_exchanger(self, config: Config) -> ParameterExchanger:
        raise NotImplementedError("Provide a partial exchanger (DynamicLayerExchanger, SparseCooParameterExchanger, ...)")

    def setup_client(self, config: Config) -> None:
        super().setup_client(config)
        if self.store_initial_model:
            self.initial_model = copy.deepcopy(self.model).to(self.device)

    def _refresh_round_start_copy(self) -> None:
        if self.initial_model is not None:
            self.initial_model.load_state_dict(self.model.state_dict(), strict=True)

    def set_parameters(self, parameters: NDArrays, config: Config, fitting_round: bool) -> None:
        super().set_parameters(parameters, config, fitting_round)
        self._refresh_round_start_copy()

    def get_parameters(self, config: Config) -> NDArrays:
        if not self.initialized:
            return self.setup_client_and_return_all_model_parameters(config)
        return self.parameter_exchanger.push_parameters(self.model, self.initial_model, config=config)
