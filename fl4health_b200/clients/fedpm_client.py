"""FedPM client (parity: ``fl4health/clients/fedpm_client.py:18-95``): converts the model to masked layers (unless the
config says it already is) and exchanges Bernoulli-sampled binary masks through ``FedPmExchanger``."""

from __future__ import annotations

from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.common.typing import Config
from fl4health_b200.model_bases.masked_layers.masked_layers_utils import convert_to_masked_model
from fl4health_b200.parameter_exchange.fedpm_exchanger import FedPmExchanger
from fl4health_b200.parameter_exchange.parameter_exchanger_base import ParameterExchanger
from fl4health_b200.utils.config import narrow_dict_type


class FedPmClient(BasicClient):
    def get_model(self, config: Config):  # noqa: ANN201
        raise NotImplementedError

    def _place_model(self, model, with_grad: bool = True):  # noqa: ANN001, ANN201
        # masks/scores are the trainable state: convert *before* the arena is laid out so scores are arena-resident
        config = getattr(self, "_setup_config", {})
        if not bool(config.get("is_masked_model", False)):
            model = convert_to_masked_model(model)
        return super()._place_model(model, with_grad)

    def setup_client(self, config: Config) -> None:
        narrow_dict_type(config, "is_masked_model", bool)
        self._setup_config = config
        super().setup_client(config)

    def get_parameter_exchanger(self, config: Config) -> ParameterExchanger:
        return FedPmExchanger()
