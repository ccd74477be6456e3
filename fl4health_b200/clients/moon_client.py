"""MOON (model-contrastive FL, Li et al. 2021).

Parity: ``fl4health/clients/moon_client.py:19-258``: frozen copies of the last ``len_old_models_buffer`` local models
and of the round-start global model give negative / positive feature views; loss = CE + mu * contrastive.  First round
(no old models yet) trains with plain CE.  The frozen networks live in a ``SnapshotBank`` and run under ``no_grad``, the
contrastive head is the fused ``MoonContrastiveLoss`` kernel, and the features of all old models are written straight
into one preallocated [N,B,F] tensor.
"""

from __future__ import annotations

from logging import WARNING
from typing import Any

import torch

from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config
from fl4health_b200.engine.snapshots import SnapshotBank
from fl4health_b200.losses.contrastive_loss import MoonContrastiveLoss
from fl4health_b200.model_bases.sequential_split_models import SequentiallySplitModel
from fl4health_b200.utils.losses import EvaluationLosses, TrainingLosses
from fl4health_b200.utils.typing import TorchFeatureType, TorchInputType, TorchPredType, TorchTargetType


class MoonClient(BasicClient):
    def __init__(self, *args: Any, temperature: float = 0.5, contrastive_weight: float = 1.0, len_old_models_buffer: int = 1,
                 **kwargs: Any) -> None:
        """Positional / keyword arguments other than the three MOON knobs are ``BasicClient``'s."""
        super().__init__(*args, **kwargs)
        if contrastive_weight == 0:
            log(WARNING, "Contrastive loss weight is set to 0, thus Contrastive loss will not be computed.")
        self.temperature, self.contrastive_weight = temperature, contrastive_weight
        self.len_old_models_buffer = len_old_models_buffer
        self.contrastive_loss_function = MoonContrastiveLoss(self.device, temperature=temperature)
        self._frozen = SnapshotBank(previous_local=len_old_models_buffer, round_start_global=1)

    # the reference's attribute names
    @property
    def old_models_list(self) -> list[torch.nn.Module]:
        return self._frozen.all("previous_local")

    @property
    def global_model(self) -> torch.nn.Module | None:
        return self._frozen.get("round_start_global")

    def _graph_variant(self) -> object:
        return self._frozen.variant()  # extra forwards appear once old models exist; their tensors change every round

    # ---------------------------------------------------------------------------------------------- round boundaries
    def update_before_train(self, current_server_round: int) -> None:
        self._frozen.capture("round_start_global", self.model)  # the model exactly as received from the server
        super().update_before_train(current_server_round)

    def update_after_train(self, local_steps: int, loss_dict: dict[str, float], config: Config) -> None:
        assert isinstance(self.model, SequentiallySplitModel)
        self._frozen.resize("previous_local", self.len_old_models_buffer)  # the attribute may be tuned after construction
        self._frozen.capture("previous_local", self.model)
        super().update_after_train(local_steps, loss_dict, config)

    # ---------------------------------------------------------------------------------------------- step
    def predict(self, input: TorchInputType) -> tuple[TorchPredType, TorchFeatureType]:
        assert isinstance(input, torch.Tensor)
        preds, features = self.model(input)
        assert "features" in features, "Model must produce a features dictionary with a 'features' key"
        live = features["features"]
        with torch.no_grad():
            negatives = self._frozen.all("previous_local")
            if negatives:
                stacked = torch.empty(len(negatives), *live.shape, dtype=live.dtype, device=self.device)
                for row, frozen in zip(stacked, negatives):
                    row.copy_(frozen(input)[1]["features"])
                features["old_features"] = stacked
            positive = self._frozen.get("round_start_global")
            if positive is not None:
                features["global_features"] = positive(input)[1]["features"]
        return preds, features

    def compute_loss_and_additional_losses(
        self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType
    ) -> tuple[torch.Tensor, dict[str, torch.Tensor]]:
        task = self.criterion(preds["prediction"], target)
        recorded = {"loss": task}
        total = task.clone()
        if {"old_features", "global_features"} <= set(features):
            contrast = self.contrastive_loss_function(features["features"], features["global_features"].unsqueeze(0),
                                                      features["old_features"])
            recorded["contrastive_loss"] = contrast
            total = total + self.contrastive_weight * contrast
        recorded["total_loss"] = total
        return total, recorded

    def _losses(self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType):  # noqa: ANN202
        """(objective, recorded dict): plain task loss until a previous local model exists."""
        contrastive = self._frozen.filled("previous_local")
        source = self if contrastive else BasicClient
        return source.compute_loss_and_additional_losses(self, preds, features, target) if not contrastive \
            else self.compute_loss_and_additional_losses(preds, features, target)

    def compute_training_loss(self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType) -> TrainingLosses:
        assert self.model.training
        objective, recorded = self._losses(preds, features, target)
        return TrainingLosses(backward=objective, additional_losses=recorded)

    def compute_evaluation_loss(self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType) -> EvaluationLosses:
        assert not self.model.training
        objective, recorded = self._losses(preds, features, target)
        checkpoint = recorded["loss"] if recorded and "loss" in recorded else objective
        return EvaluationLosses(checkpoint=checkpoint, additional_losses=recorded)
