"""Ditto (Li et al. 2021): a global model trained FedAvg-style plus a personal model regularised towards it.

Parity: ``fl4health/clients/ditto_client.py:20-400``: two models (``global_model`` exchanged with the server,
``model`` personal), two optimizers (``"global"``, ``"local"``), two backward passes per step, penalty
``lambda/2 ||w_personal - w_global_init||^2``; predictions keyed ``"global"`` / ``"local"``.
"""

from __future__ import annotations

from collections.abc import Sequence
from logging import INFO
from pathlib import Path

import torch
from torch import nn
from torch.optim import Optimizer

from fl4health_b200.checkpointing.client_module import ClientCheckpointAndStateModule
from fl4health_b200.clients.adaptive_drift_constraint_client import AdaptiveDriftConstraintClient
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config, NDArrays, Scalar
from fl4health_b200.engine.options import EngineOptions
from fl4health_b200.metrics.base_metrics import Metric
from fl4health_b200.reporting.base_reporter import BaseReporter
from fl4health_b200.utils.config import narrow_dict_type
from fl4health_b200.utils.losses import EvaluationLosses, LossMeterType, TrainingLosses
from fl4health_b200.utils.typing import TorchFeatureType, TorchInputType, TorchPredType, TorchTargetType
from fl4health_b200.parameter_exchange.full_exchanger import FullParameterExchanger


class DittoClient(AdaptiveDriftConstraintClient):
    def __init__(
        self,
        data_path: Path,
        metrics: Sequence[Metric],
        device: torch.device,
        loss_meter_type: LossMeterType = LossMeterType.AVERAGE,
        checkpoint_and_state_module: ClientCheckpointAndStateModule | None = None,
        reporters: Sequence[BaseReporter] | None = None,
        progress_bar: bool = False,
        client_name: str | None = None,
        engine_options: EngineOptions | None = None,
    ) -> None:
        super().__init__(
            data_path=data_path, metrics=metrics, device=device, loss_meter_type=loss_meter_type,
            checkpoint_and_state_module=checkpoint_and_state_module, reporters=reporters, progress_bar=progress_bar,
            client_name=client_name, engine_options=engine_options,
        )
        self.global_model: nn.Module
        self.penalty_optimizer_key = "local"

    def get_optimizer(self, config: Config) -> dict[str, Optimizer]:
        raise NotImplementedError(
            "User Clients must define a function that returns a dict[str, Optimizer] with keys 'global' and 'local' "
            "defining separate optimizers for the global and local models of Ditto."
        )

    def set_optimizer(self, config: Config) -> None:
        optimizers = self.get_optimizer(config)
        assert isinstance(optimizers, dict) and set(optimizers.keys()) == {"global", "local"}
        self.optimizers = optimizers

    def get_global_model(self, config: Config) -> nn.Module:
        """Architecture of the global model (defaults to the personal model's architecture)."""
        return self.get_model(config)

    def _candidate_modules(self) -> list[nn.Module]:
        return [self.model, self.global_model]

    def setup_client(self, config: Config) -> None:
        self.global_model = self._place_model(self.get_global_model(config))
        super().setup_client(config)

    # ------------------------------------------------------------------------------------------ exchange
    def get_parameters(self, config: Config) -> NDArrays:
        if not self.initialized:
            return self.setup_client_and_return_all_model_parameters(config)
        if self.initial_parameters_requested(config):  # already set up by a properties poll: plain model state, unpacked
            return FullParameterExchanger().push_parameters(self.model, config=config)
        assert self.global_model is not None and self.parameter_exchanger is not None
        weights = self.parameter_exchanger.push_parameters(self.global_model, config=config)
        return self.parameter_exchanger.pack_parameters(weights, self.loss_for_adaptation)

    def set_parameters(self, parameters: NDArrays, config: Config, fitting_round: bool) -> None:
        """Server weights go to the GLOBAL model; on round 1 they also initialise the personal model."""
        assert self.global_model is not None and self.model is not None and self.parameter_exchanger is not None
        server_model_state, self.drift_penalty_weight = self.parameter_exchanger.unpack_parameters(parameters)
        log(INFO, f"Lambda weight received from the server: {self.drift_penalty_weight}")
        current_server_round = narrow_dict_type(config, "current_server_round", int)
        if current_server_round == 1 and fitting_round:
            log(INFO, "Initializing the global and local models weights for the first time")
            self.initialize_all_model_weights(server_model_state, config)
        else:
            self.parameter_exchanger.pull_parameters(server_model_state, self.global_model, config)

    def initialize_all_model_weights(self, parameters: NDArrays, config: Config) -> None:
        self.parameter_exchanger.pull_parameters(parameters, self.model, config)
        self.parameter_exchanger.pull_parameters(parameters, self.global_model, config)

    def set_initial_global_tensors(self) -> None:
        self.drift_penalty_tensors = self.snapshot_drift_anchor(source_model=self.global_model, constrained_model=self.model)

    def update_before_train(self, current_server_round: int) -> None:
        self.set_initial_global_tensors()
        self.global_model.train()
        super().update_before_train(current_server_round)

    # ------------------------------------------------------------------------------------------ step
    def train_step(self, input: TorchInputType, target: TorchTargetType) -> tuple[TrainingLosses, TorchPredType]:
        self.optimizers["global"].zero_grad()
        self.optimizers["local"].zero_grad()
        with self._amp():
            preds, features = self.predict(input)
            target = self.transform_target(target)
            losses = self.compute_training_loss(preds, features, target)
        losses.additional_losses["global_loss"].backward()
        self.optimizers["global"].step()
        losses.backward["backward"].backward()
        self.optimizers["local"].step()
        return losses, preds

    def predict(self, input: TorchInputType) -> tuple[TorchPredType, TorchFeatureType]:
        if isinstance(input, torch.Tensor):
            global_preds, local_preds = self.global_model(input), self.model(input)
        elif isinstance(input, dict):
            global_preds, local_preds = self.global_model(**input), self.model(**input)
        else:
            raise TypeError('"input" must be of type torch.Tensor or dict[str, torch.Tensor].')
        assert isinstance(global_preds, torch.Tensor) and isinstance(local_preds, torch.Tensor)
        return {"global": global_preds, "local": local_preds}, {}

    def compute_loss_and_additional_losses(
        self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType
    ) -> tuple[torch.Tensor, dict[str, torch.Tensor]]:
        global_loss = self.criterion(preds["global"], target)
        local_loss = self.criterion(preds["local"], target)
        return local_loss, {"local_loss": local_loss.clone(), "global_loss": global_loss}

    def compute_training_loss(
        self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType
    ) -> TrainingLosses:
        assert self.global_model.training and self.model.training
        loss, additional_losses = self.compute_loss_and_additional_losses(preds, features, target)
        additional_losses["loss_for_adaptation"] = additional_losses["local_loss"].clone()
        penalty_loss = self.compute_penalty_loss()
        additional_losses["penalty_loss"] = penalty_loss.clone()
        return TrainingLosses(backward=loss + penalty_loss, additional_losses=additional_losses)

    def validate(self, include_losses_in_metrics: bool = False) -> tuple[float, dict[str, Scalar]]:
        self.global_model.eval()
        return super().validate(include_losses_in_metrics=include_losses_in_metrics)

    def compute_evaluation_loss(
        self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType
    ) -> EvaluationLosses:
        assert not self.global_model.training and not self.model.training
        return super().compute_evaluation_loss(preds, features, target)
