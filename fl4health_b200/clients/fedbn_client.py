"""FedBN: everything except normalisation layers is exchanged (parity: ``fedbn_client.py:7-28``)."""

from __future__ import annotations

from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.common.typing import Config
from fl4health_b200.parameter_exchange.layer_exchanger import LayerExchangerWithExclusions
from fl4health_b200.parameter_exchange.parameter_exchanger_base import ParameterExchanger


class FedBnClient(BasicClient):
    def get_parameter_exchanger(self, config: Config) -> ParameterExchanger:
        raise NotImplementedError(
            "FedBnClient needs a LayerExchangerWithExclusions, e.g. "
            "LayerExchangerWithExclusions(self.model, {nn.BatchNorm2d}); override get_parameter_exchanger."
        )

    def setup_client(self, config: Config) -> None:
        super().setup_client(config)
        assert isinstance(self.parameter_exchanger, LayerExchangerWithExclusions), (
            "FedBN requires a LayerExchangerWithExclusions parameter exchanger"
        )
