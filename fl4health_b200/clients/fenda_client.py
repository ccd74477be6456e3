"""FENDA-FL client: parallel local/global extractors, only the global extractor is exchanged
(parity: ``fl4health/clients/fenda_client.py:17-70``)."""

from __future__ import annotations

from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.common.typing import Config
from fl4health_b200.model_bases.fenda_base import FendaModel
from fl4health_b200.parameter_exchange.layer_exchanger import FixedLayerExchanger
from fl4health_b200.parameter_exchange.parameter_exchanger_base import ParameterExchanger


class FendaClient(BasicClient):
    def get_parameter_exchanger(self, config: Config) -> ParameterExchanger:
        assert isinstance(self.model, FendaModel)
        return FixedLayerExchanger(self.model.layers_to_exchange())
