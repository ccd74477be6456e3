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
    def __init__(self, *args: Any, store_initial_model: bool = False, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self.initial_model: nn.Module | None = None
        self.store_initial_model = store_initial_model

    def setup_client(self, config: Config) -> None:
        super().setup_client(config)
        self.initial_model = copy.deepcopy(self.model).to(self.device) if self.store_initial_model else None

    def get_parameter_exchanger(self, config: Config) -> ParameterExchanger:
        raise NotImplementedError("Provide a partial exchanger (DynamicLayerExchanger, SparseCooParameterExchanger, ...)")

    def get_parameters(self, config: Config) -> NDArrays:
        if not self.initialized:
            return self.setup_client_and_return_all_model_parameters(config)
        assert self.model is not None and self.parameter_exchanger is not None
        return self.parameter_exchanger.push_parameters(self.model, self.initial_model, config=config)

    def set_parameters(self, parameters: NDArrays, config: Config, fitting_round: bool) -> None:
        super().set_parameters(parameters, config, fitting_round)
        if self.store_initial_model:
            assert self.initial_model is not None
            self.initial_model.load_state_dict(self.model.state_dict(), strict=True)
