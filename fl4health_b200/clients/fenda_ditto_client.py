"""FENDA + Ditto (parity: ``fl4health/clients/fenda_ditto_client.py:21-345``): the personal model is a FENDA model whose
global extractor is refreshed each round from a Ditto-style global ``SequentiallySplitModel`` feature extractor; the
drift penalty ties a FENDA extractor to that global extractor."""

from __future__ import annotations

from collections.abc import Sequence
from logging import INFO
from pathlib import Path
from typing import Any

import torch

from fl4health_b200.clients.ditto_client import DittoClient
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config, NDArrays
from fl4health_b200.metrics.base_metrics import Metric
from fl4health_b200.model_bases.fenda_base import FendaModel
from fl4health_b200.model_bases.sequential_split_models import SequentiallySplitModel
from fl4health_b200.parameter_exchange.packing_exchanger import FullParameterExchangerWithPacking
from fl4health_b200.utils.losses import LossMeterType, TrainingLosses
from fl4health_b200.utils.typing import TorchFeatureType, TorchInputType, TorchPredType, TorchTargetType


def check_shape_match(params1: Any, params2: Any, error_message: str) -> None:
    p1, p2 = list(params1), list(params2)
    assert len(p1) == len(p2) and all(a.shape == b.shape for a, b in zip(p1, p2)), error_message


class FendaDittoClient(DittoClient):
    def __init__(
        self,
        data_path: Path,
        metrics: Sequence[Metric],
        device: torch.device,
        loss_meter_type: LossMeterType = LossMeterType.AVERAGE,
        checkpoint_and_state_module: Any = None,
        reporters: Any = None,
        progress_bar: bool = False,
        client_name: str | None = None,
        freeze_global_feature_extractor: bool = False,
        engine_options: Any = None,
    ) -> None:
        super().__init__(data_path=data_path, metrics=metrics, device=device, loss_meter_type=loss_meter_type,
                         checkpoint_and_state_module=checkpoint_and_state_module, reporters=reporters,
                         progress_bar=progress_bar, client_name=client_name, engine_options=engine_options)
        self.global_model: SequentiallySplitModel
        self.model: FendaModel
        self.freeze_global_feature_extractor = freeze_global_feature_extractor

    def get_model(self, config: Config) -> FendaModel:
        raise NotImplementedError("This function must be defined in the inheriting class to use this client")

    def get_global_model(self, config: Config) -> SequentiallySplitModel:
        raise NotImplementedError("This function must be defined in the inheriting class to use this client")

    def _check_shape_match(self) -> None:
        check_shape_match(self.global_model.base_module.parameters(), self.model.second_feature_extractor.parameters(),
                          "global_model.base_module and model.second_feature_extractor must match exactly.")
        check_shape_match(self.model.second_feature_extractor.parameters(), self.model.first_feature_extractor.parameters(),
                          "model.second_feature_extractor and model.first_feature_extractor must match exactly.")

    def setup_client(self, config: Config) -> None:
        super().setup_client(config)
        self._check_shape_match()

    def _fused_penalty_optimizer(self) -> None:  # the constrained tensors are a sub-module: use the autograd penalty
        return None

    def get_parameters(self, config: Config) -> NDArrays:
        if not self.initialized:
            # server-side initialisation request: hand over the GLOBAL model's weights; the strategy appends the
            # penalty weight itself (``add_auxiliary_information``)
            log(INFO, "Setting up client")
            self.setup_client(config)
            return self.parameter_exchanger.push_parameters(self.global_model, config=config)
        weights = self.parameter_exchanger.push_parameters(self.global_model, config=config)
        return self.parameter_exchanger.pack_parameters(weights, self.loss_for_adaptation)

    def set_parameters(self, parameters: NDArrays, config: Config, fitting_round: bool) -> None:
        assert isinstance(self.parameter_exchanger, FullParameterExchangerWithPacking)
        server_model_state, self.drift_penalty_weight = self.parameter_exchanger.unpack_parameters(parameters)
        log(INFO, f"Penalty weight received from the server: {self.drift_penalty_weight}")
        self.parameter_exchanger.pull_parameters(server_model_state, self.global_model, config)
        self.model.second_feature_extractor.load_state_dict(self.global_model.base_module.state_dict())

    def set_initial_global_tensors(self) -> None:
        self.drift_penalty_tensors = [p.detach().clone() for p in self.global_model.base_module.parameters()]

    def update_before_train(self, current_server_round: int) -> None:
        if self.freeze_global_feature_extractor:
            for param in self.model.second_feature_extractor.parameters():
                param.requires_grad = False
        return super().update_before_train(current_server_round)

    def predict(self, input: TorchInputType) -> tuple[TorchPredType, TorchFeatureType]:
        if isinstance(input, torch.Tensor):
            global_preds, _ = self.global_model(input)
            local_preds, _ = self.model(input)
        else:
            global_preds, _ = self.global_model(**input)
            local_preds, _ = self.model(**input)
        return {"global": global_preds["prediction"], "local": local_preds["prediction"]}, {}

    def compute_training_loss(self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType) -> TrainingLosses:
        assert self.global_model.training and self.model.training
        loss, additional = self.compute_loss_and_additional_losses(preds, features, target)
        additional = additional or {}
        additional["loss_for_adaptation"] = loss.clone()
        constrained = self.model.first_feature_extractor if self.freeze_global_feature_extractor else self.model.second_feature_extractor
        assert self.drift_penalty_tensors is not None and self.drift_penalty_weight is not None
        penalty = self.penalty_loss_function(constrained, self.drift_penalty_tensors, self.drift_penalty_weight)
        additional["penalty_loss"] = penalty.clone()
        return TrainingLosses(backward=loss + penalty, additional_losses=additional)
