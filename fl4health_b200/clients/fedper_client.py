"""FedPer: exchange only the base (feature extractor), keep the head personal (parity: ``fedper_client.py:9-24``)."""

from __future__ import annotations

from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.common.typing import Config
from fl4health_b200.model_bases.sequential_split_models import SequentiallySplitExchangeBaseModel
from fl4health_b200.parameter_exchange.layer_exchanger import FixedLayerExchanger
from fl4health_b200.parameter_exchange.parameter_exchanger_base import ParameterExchanger


class FedPerClient(BasicClient):
    def get_parameter_exchanger(self, config: Config) -> ParameterExchanger:
        assert isinstance(self.model, SequentiallySplitExchangeBaseModel)
        return FixedLayerExchanger(self.model.layers_to_exchange())
