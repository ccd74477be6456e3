"""FENDA + Ditto (parity: ``fl4health/clients/fenda_ditto_client.py:21-345``): the personal model is a FENDA model whose
global extractor is refreshed each round from a Ditto-style global ``SequentiallySplitModel`` feature extractor; the
drift penalty ties a FENDA extractor to that global extractor."""

from __future__ import annotations

from logging import INFO
from typing import Any

import torch

from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.clients.ditto_client import DittoClient
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config, NDArrays
from fl4health_b200.engine import outputs as model_outputs
from fl4health_b200.model_bases.fenda_base import FendaModel
from fl4health_b200.model_bases.sequential_split_models import SequentiallySplitModel
from fl4health_b200.parameter_exchange.packing_exchanger import FullParameterExchangerWithPacking
from fl4health_b200.utils.losses import TrainingLosses
from fl4health_b200.utils.typing import TorchFeatureType, TorchInputType, TorchPredType, TorchTargetType


def check_shape_match(params1: Any, params2: Any, error_message: str) -> None:
    shapes = [[tuple(p.shape) for p in group] for group in (params1, params2)]
    assert shapes[0] == shapes[1], error_message


class FendaDittoClient(DittoClient):
    """Ditto whose personal network is a FENDA model: every round the server's aggregate lands in the global twin (a
    ``SequentiallySplitModel``) and its feature extractor is copied into FENDA's *second* extractor; the drift penalty
    ties a FENDA extractor (the second one, or the first when the second is frozen) to that received extractor."""

    model: FendaModel
    global_model: SequentiallySplitModel

    def __init__(self, *args: Any, freeze_global_feature_extractor: bool = False, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self.freeze_global_feature_extractor = freeze_global_feature_extractor

    def get_model(self, config: Config) -> FendaModel:
        raise NotImplementedError("This function must be defined in the inheriting class to use this client")

    def get_global_model(self, config: Config) -> SequentiallySplitModel:
        raise NotImplementedError("This function must be defined in the inheriting class to use this client")

    def _fused_penalty_optimizer(self) -> None:  # the constrained tensors are a sub-module: use the autograd penalty
        return None

    def _check_shape_match(self) -> None:
        received, second, first = (self.global_model.base_module, self.model.second_feature_extractor,
                                   self.model.first_feature_extractor)
        check_shape_match(received.parameters(), second.parameters(),
                          "global_model.base_module and model.second_feature_extractor must match exactly.")
        check_shape_match(second.parameters(), first.parameters(),
                          "model.second_feature_extractor and model.first_feature_extractor must match exactly.")

    def setup_client(self, config: Config) -> None:
        super().setup_client(config)
        self._check_shape_match()

    # ------------------------------------------------------------------------------------------ exchange
    def get_parameters(self, config: Config) -> NDArrays:
        first_contact = not self.initialized
        if first_contact:
            log(INFO, "Setting up client")
            self.setup_client(config)
        outgoing = self.parameter_exchanger.push_parameters(self.global_model, config=config)
        # initialisation request: bare global weights (the strategy appends the penalty weight itself)
        return outgoing if first_contact else self.parameter_exchanger.pack_parameters(outgoing, self.loss_for_adaptation)

    def _install_aggregate(self, aggregate: NDArrays, config: Config, fitting_round: bool) -> None:
        assert isinstance(self.parameter_exchanger, FullParameterExchangerWithPacking)
        self.parameter_exchanger.pull_parameters(aggregate, self.global_model, config)
        self.model.second_feature_extractor.load_state_dict(self.global_model.base_module.state_dict())

    def set_initial_global_tensors(self) -> None:
        self.drift_penalty_tensors = [p.detach().clone() for p in self.global_model.base_module.parameters()]

    def update_before_train(self, current_server_round: int) -> None:
        if self.freeze_global_feature_extractor:
            self.model.second_feature_extractor.requires_grad_(False)
        self.set_initial_global_tensors()
        BasicClient.update_before_train(self, current_server_round)

    # ------------------------------------------------------------------------------------------ step
    def predict(self, input: TorchInputType) -> tuple[TorchPredType, TorchFeatureType]:
        heads = {role: model_outputs.call_model(net, input)[0]["prediction"] for role, net in self._networks().items()}
        return heads, {}

    def _constrained_extractor(self) -> torch.nn.Module:
        return self.model.first_feature_extractor if self.freeze_global_feature_extractor else self.model.second_feature_extractor

    def compute_training_loss(self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType) -> TrainingLosses:
        assert self.global_model.training and self.model.training
        assert self.drift_penalty_tensors is not None and self.drift_penalty_weight is not None
        personal, recorded = self.compute_loss_and_additional_losses(preds, features, target)
        penalty = self.penalty_loss_function(self._constrained_extractor(), self.drift_penalty_tensors, self.drift_penalty_weight)
        recorded = {**(recorded or {}), "loss_for_adaptation": personal.clone(), "penalty_loss": penalty.clone()}
        return TrainingLosses(backward=personal + penalty, additional_losses=recorded)
