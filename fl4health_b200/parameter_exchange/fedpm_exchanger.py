"""FedPM exchange: push sampled binary masks, pull logit(theta) into the score tensors
(parity: ``fl4health/parameter_exchange/fedpm_exchanger.py:10-26``)."""

from __future__ import annotations

import torch
from torch import nn

from fl4health_b200.common.typing import Config, NDArrays, to_tensor
from fl4health_b200.parameter_exchange._state import inject_state
from fl4health_b200.parameter_exchange.layer_exchanger import DynamicLayerExchanger
from fl4health_b200.parameter_exchange.parameter_selection_criteria import select_scores_and_sample_masks
from fl4health_b200.utils.functions import sigmoid_inverse


class FedPmExchanger(DynamicLayerExchanger):
    def __init__(self) -> None:
        super().__init__(select_scores_and_sample_masks)

    def pull_parameters(self, parameters: NDArrays, model: nn.Module, config: Config | None = None) -> None:
        probabilities, names = self.unpack_parameters(parameters)
        with torch.no_grad():
            scores = [sigmoid_inverse(to_tensor(p).float()) for p in probabilities]
        inject_state(model, names, scores)
