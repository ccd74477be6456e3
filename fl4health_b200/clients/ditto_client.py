"""Ditto (Li et al. 2021): a global model trained FedAvg-style plus a personal model regularised towards it.

Parity: ``fl4health/clients/ditto_client.py:20-400``: two models (``global_model`` exchanged with the server,
``model`` personal), two optimizers (``"global"``, ``"local"``), two backward passes per step, penalty
``lambda/2 ||w_personal - w_global_init||^2``; predictions keyed ``"global"`` / ``"local"``.

Declared on top of ``AdaptiveDriftConstraintClient``: the global twin is a trainable *companion* (built, placed, mode-
switched and offered to the optimizer translation by the engine), it is the exchanged network and — as received at the
start of the round — the anchor of the personal model, whose ``"local"`` optimizer absorbs the penalty gradient.
"""

from __future__ import annotations

from typing import Any

import torch
from torch import nn
from torch.optim import Optimizer

from fl4health_b200.clients.adaptive_drift_constraint_client import AdaptiveDriftConstraintClient
from fl4health_b200.common.typing import Config, NDArrays
from fl4health_b200.engine.companions import FOLLOW, Companion
from fl4health_b200.utils.losses import TrainingLosses
from fl4health_b200.utils.typing import TorchFeatureType, TorchInputType, TorchPredType, TorchTargetType

_ROLES = ("global", "local")  # optimizer keys == prediction keys


class DittoClient(AdaptiveDriftConstraintClient):
    companions = {"global_model": Companion(factory="get_global_model", trainable=True, mode=FOLLOW)}
    exchanged_model = "global_model"
    receives_into = "global_model"
    anchor_model = "global_model"
    penalty_optimizer_key = "local"

    global_model: nn.Module

    # ------------------------------------------------------------------------------------------ user factories
    def get_global_model(self, config: Config) -> nn.Module:
        """Architecture of the global model (defaults to the personal model's architecture)."""
        return self.get_model(config)

    def get_optimizer(self, config: Config) -> dict[str, Optimizer]:
        raise NotImplementedError(
            "User Clients must define a function that returns a dict[str, Optimizer] with keys 'global' and 'local' "
            "defining separate optimizers for the global and local models of Ditto."
        )

    def set_optimizer(self, config: Config) -> None:
        optimizers = self.get_optimizer(config)
        assert isinstance(optimizers, dict) and set(optimizers) == set(_ROLES)
        self.optimizers = optimizers

    def initialize_all_model_weights(self, parameters: NDArrays, config: Config) -> None:
        for network in (self.model, self.global_model):  # round 1: both twins start from the server's weights
            self.parameter_exchanger.pull_parameters(parameters, network, config)

    def set_initial_global_tensors(self) -> None:
        self.drift_penalty_tensors = self.snapshot_drift_anchor(source_model=self.global_model, constrained_model=self.model)

    # ------------------------------------------------------------------------------------------ step
    def _networks(self) -> dict[str, nn.Module]:
        return {"global": self.global_model, "local": self.model}

    def predict(self, input: TorchInputType) -> tuple[TorchPredType, TorchFeatureType]:
        if not isinstance(input, (torch.Tensor, dict)):
            raise TypeError('"input" must be of type torch.Tensor or dict[str, torch.Tensor].')
        outputs: dict[str, Any] = {role: (net(**input) if isinstance(input, dict) else net(input))
                                   for role, net in self._networks().items()}
        assert all(isinstance(out, torch.Tensor) for out in outputs.values())
        return outputs, {}

    def compute_loss_and_additional_losses(
        self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType
    ) -> tuple[torch.Tensor, dict[str, torch.Tensor]]:
        per_role = {role: self.criterion(preds[role], target) for role in _ROLES}
        return per_role["local"], {"local_loss": per_role["local"].clone(), "global_loss": per_role["global"]}

    def compute_training_loss(self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType) -> TrainingLosses:
        personal, recorded = self.compute_loss_and_additional_losses(preds, features, target)
        penalty = self.compute_penalty_loss()
        recorded.update(loss_for_adaptation=recorded["local_loss"].clone(), penalty_loss=penalty.clone())
        return TrainingLosses(backward=personal + penalty, additional_losses=recorded)

    def train_step(self, input: TorchInputType, target: TorchTargetType) -> tuple[TrainingLosses, TorchPredType]:
        """Two independent updates per batch: the global twin on its plain loss, the personal model on loss + penalty."""
        for role in _ROLES:
            self.optimizers[role].zero_grad()
        with self._amp():
            preds, features = self.predict(input)
            losses = self.compute_training_loss(preds, features, self.transform_target(target))
        for role, objective in (("global", losses.additional_losses["global_loss"]), ("local", losses.backward["backward"])):
            objective.backward()
            self.optimizers[role].step()
        return losses, preds
