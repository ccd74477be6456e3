"""Ensemble client (parity: ``fl4health/clients/ensemble_client.py:17-196``): one optimizer per ensemble member, a dict
of backward losses (one per member), checkpoint loss on the ensemble prediction."""

from __future__ import annotations

import torch
from torch.optim import Optimizer

from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.common.typing import Config
from fl4health_b200.model_bases.ensemble_base import EnsembleModel
from fl4health_b200.utils.losses import EvaluationLosses, TrainingLosses
from fl4health_b200.utils.typing import TorchFeatureType, TorchInputType, TorchPredType, TorchTargetType


class EnsembleClient(BasicClient):
    model: EnsembleModel

    def setup_client(self, config: Config) -> None:
        super().setup_client(config)
        assert len(self.optimizers) == len(self.model.ensemble_models)
        assert sorted(self.optimizers.keys()) == sorted(self.model.ensemble_models.keys())

    def set_optimizer(self, config: Config) -> None:
        optimizers = self.get_optimizer(config)
        assert isinstance(optimizers, dict)
        self.optimizers = optimizers

    def train_step(self, input: TorchInputType, target: TorchTargetType) -> tuple[TrainingLosses, TorchPredType]:
        assert isinstance(input, torch.Tensor)
        for optimizer in self.optimizers.values():
            optimizer.zero_grad()
        with self._amp():
            preds, features = self.predict(input)
            target = self.transform_target(target)
            losses = self.compute_training_loss(preds, features, target)
        for loss in losses.backward.values():
            loss.backward()
        for optimizer in self.optimizers.values():
            optimizer.step()
        return losses, preds

    def compute_training_loss(self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType) -> TrainingLosses:
        losses = {key: self.criterion(pred.float(), target) for key, pred in preds.items() if key != "ensemble-pred"}
        return TrainingLosses(backward=losses)

    def compute_evaluation_loss(self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType) -> EvaluationLosses:
        return EvaluationLosses(checkpoint=self.criterion(preds["ensemble-pred"].float(), target))

    def get_optimizer(self, config: Config) -> dict[str, Optimizer]:
        raise NotImplementedError("Return one optimizer per ensemble member, keyed like EnsembleModel.ensemble_models")
