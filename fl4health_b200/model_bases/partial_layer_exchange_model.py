"""Models that expose the subset of state they exchange (parity: ``partial_layer_exchange_model.py:6-9``)."""

from abc import ABC, abstractmethod

from torch import nn


class PartialLayerExchangeModel(nn.Module, ABC):
    @abstractmethod
    def layers_to_exchange(self) -> list[str]:
        raise NotImplementedError
