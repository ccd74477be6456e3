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

_VOTE = "ensemble-pred"  # key of the combined prediction; every other key is one member


class EnsembleClient(BasicClient):
    model: EnsembleModel

    def get_optimizer(self, config: Config) -> dict[str, Optimizer]:
        raise NotImplementedError("Return one optimizer per ensemble member, keyed like EnsembleModel.ensemble_models")

    def set_optimizer(self, config: Config) -> None:
        per_member = self.get_optimizer(config)
        assert isinstance(per_member, dict)
        self.optimizers = per_member

    def setup_client(self, config: Config) -> None:
        super().setup_client(config)
        assert set(self.optimizers) == set(self.model.ensemble_models) and len(self.optimizers) == len(self.model.ensemble_models)

    def compute_training_loss(self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType) -> TrainingLosses:
        """Members are trained independently: one backward loss per member, nothing on the vote."""
        return TrainingLosses(backward={member: self.criterion(logits.float(), target)
                                        for member, logits in preds.items() if member != _VOTE})

    def compute_evaluation_loss(self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType) -> EvaluationLosses:
        return EvaluationLosses(checkpoint=self.criterion(preds[_VOTE].float(), target))

    def train_step(self, input: TorchInputType, target: TorchTargetType) -> tuple[TrainingLosses, TorchPredType]:
        assert isinstance(input, torch.Tensor)
        members = list(self.optimizers.values())
        for optimizer in members:
            optimizer.zero_grad()
        with self._amp():
            preds, features = self.predict(input)
            losses = self.compute_training_loss(preds, features, self.transform_target(target))
        for member_loss in losses.backward.values():  # the members share no parameters: order is irrelevant
            member_loss.backward()
        for optimizer in members:
            optimizer.step()
        return losses, preds
