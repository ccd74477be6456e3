"""MOON (model-contrastive FL, Li et al. 2021).

Parity: ``fl4health/clients/moon_client.py:19-258``: frozen copies of the last ``len_old_models_buffer`` local models
and of the round-start global model give negative / positive feature views; loss = CE + mu * contrastive.  First round
(no old models yet) trains with plain CE.  The frozen models run in inference mode (no autograd state kept), the
contrastive head is the fused ``MoonContrastiveLoss`` kernel, and the features of all old models are written straight
into one preallocated [N,B,F] tensor.
"""

from __future__ import annotations

from collections.abc import Sequence
from logging import WARNING
from pathlib import Path

import torch

from fl4health_b200.checkpointing.client_module import ClientCheckpointAndStateModule
from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config
from fl4health_b200.engine.options import EngineOptions
from fl4health_b200.losses.contrastive_loss import MoonContrastiveLoss
from fl4health_b200.metrics.base_metrics import Metric
from fl4health_b200.model_bases.sequential_split_models import SequentiallySplitModel
from fl4health_b200.reporting.base_reporter import BaseReporter
from fl4health_b200.utils.client import clone_and_freeze_model
from fl4health_b200.utils.losses import EvaluationLosses, LossMeterType, TrainingLosses
from fl4health_b200.utils.typing import TorchFeatureType, TorchInputType, TorchPredType, TorchTargetType


class MoonClient(BasicClient):
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
        temperature: float = 0.5,
        contrastive_weight: float = 1.0,
        len_old_models_buffer: int = 1,
        engine_options: EngineOptions | None = None,
    ) -> None:
        super().__init__(
            data_path=data_path, metrics=metrics, device=device, loss_meter_type=loss_meter_type,
            checkpoint_and_state_module=checkpoint_and_state_module, reporters=reporters, progress_bar=progress_bar,
            client_name=client_name, engine_options=engine_options,
        )
        self.temperature = temperature
        self.contrastive_weight = contrastive_weight
        if contrastive_weight == 0:
            log(WARNING, "Contrastive loss weight is set to 0, thus Contrastive loss will not be computed.")
        self.contrastive_loss_function = MoonContrastiveLoss(self.device, temperature=temperature)
        self.len_old_models_buffer = len_old_models_buffer
        self.old_models_list: list[torch.nn.Module] = []
        self.global_model: torch.nn.Module | None = None

    def _graph_variant(self) -> object:
        # the step launches extra forwards once old models exist; frozen-model tensors are rebound every round
        return (len(self.old_models_list), id(self.global_model), tuple(id(m) for m in self.old_models_list))

    def predict(self, input: TorchInputType) -> tuple[TorchPredType, TorchFeatureType]:
        assert isinstance(input, torch.Tensor)
        preds, features = self.model(input)
        assert "features" in features, "Model must produce a features dictionary with a 'features' key"
        with torch.no_grad():
            if self.old_models_list:
                old_features = torch.empty(len(self.old_models_list), *features["features"].shape,
                                           dtype=features["features"].dtype, device=self.device)
                for idx, old_model in enumerate(self.old_models_list):
                    old_features[idx] = old_model(input)[1]["features"]
                features["old_features"] = old_features
            if self.global_model is not None:
                features["global_features"] = self.global_model(input)[1]["features"]
        return preds, features

    def update_after_train(self, local_steps: int, loss_dict: dict[str, float], config: Config) -> None:
        assert isinstance(self.model, SequentiallySplitModel)
        self.old_models_list.append(clone_and_freeze_model(self.model))
        if len(self.old_models_list) > self.len_old_models_buffer:
            self.old_models_list.pop(0)
        super().update_after_train(local_steps, loss_dict, config)

    def update_before_train(self, current_server_round: int) -> None:
        self.global_model = clone_and_freeze_model(self.model)  # the model as just received from the server
        super().update_before_train(current_server_round)

    def compute_loss_and_additional_losses(
        self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType
    ) -> tuple[torch.Tensor, dict[str, torch.Tensor]]:
        loss = self.criterion(preds["prediction"], target)
        total_loss = loss.clone()
        additional_losses = {"loss": loss}
        if "old_features" in features and "global_features" in features:
            contrastive_loss = self.contrastive_loss_function(
                features["features"], features["global_features"].unsqueeze(0), features["old_features"]
            )
            total_loss = total_loss + self.contrastive_weight * contrastive_loss
            additional_losses["contrastive_loss"] = contrastive_loss
        additional_losses["total_loss"] = total_loss
        return total_loss, additional_losses

    def compute_training_loss(self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType) -> TrainingLosses:
        assert self.model.training
        if not self.old_models_list:
            total_loss, additional_losses = BasicClient.compute_loss_and_additional_losses(self, preds, features, target)
        else:
            total_loss, additional_losses = self.compute_loss_and_additional_losses(preds, features, target)
        return TrainingLosses(backward=total_loss, additional_losses=additional_losses)

    def compute_evaluation_loss(self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType) -> EvaluationLosses:
        assert not self.model.training
        if not self.old_models_list:
            checkpoint_loss, additional_losses = BasicClient.compute_loss_and_additional_losses(self, preds, features, target)
        else:
            _, additional_losses = self.compute_loss_and_additional_losses(preds, features, target)
            checkpoint_loss = additional_losses["loss"]
        return EvaluationLosses(checkpoint=checkpoint_loss, additional_losses=additional_losses)
