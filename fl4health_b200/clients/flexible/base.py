"""FlexibleClient: the model/optimizer-parameterised client API (parity: ``fl4health/clients/flexible/base.py:28-341``).

``train_step`` / ``val_step`` / ``predict`` are expressed through helpers that take the model (and optimizer) as
arguments, so mixins can drive several models through the same user hooks (Ditto: global + personal model).  Subclasses
customise ``predict_with_model`` / ``_val_step_with_model`` / ``_train_step_with_model_and_optimizer`` (and its two
halves); overriding the legacy un-parameterised methods triggers a warning.
"""

from __future__ import annotations

import warnings
from logging import WARNING
from typing import Any

import torch
from torch import nn
from torch.optim import Optimizer

from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.common.logger import log
from fl4health_b200.engine import outputs as model_outputs
from fl4health_b200.utils.losses import EvaluationLosses, TrainingLosses
from fl4health_b200.utils.typing import TorchFeatureType, TorchInputType, TorchPredType, TorchTargetType

EXPECTED_OUTPUT_TUPLE_SIZE = model_outputs.PAIR
_LEGACY_HOOKS = {
    "predict": "predict_with_model()",
    "val_step": "_val_step_with_model()",
    "train_step": "_train_step_with_model_and_optimizer() and its helper methods",
}


class FlexibleClient(BasicClient):
    def __init_subclass__(cls, **kwargs: Any) -> None:
        super().__init_subclass__(**kwargs)
        exempt = cls.__dict__.get("_dynamically_created", False) or any(
            getattr(ancestor, "_is_flexible_mixin", False) for ancestor in cls.__mro__[1:] if ancestor is not FlexibleClient
        )
        if exempt:  # mixins legitimately re-define the legacy entry points on top of the helpers
            return
        for legacy in sorted(set(_LEGACY_HOOKS) & set(cls.__dict__)):
            message = (f"`{cls.__name__}` overrides `{legacy}()`, but this method should no longer be overridden. "
                       f"Please use `{_LEGACY_HOOKS[legacy]}` instead.")
            log(WARNING, message)
            warnings.warn(message, RuntimeWarning, stacklevel=2)

    # ---- the model-parameterised helpers (what subclasses and mixins customise) -------------------------------
    def predict_with_model(self, model: nn.Module, input: TorchInputType) -> tuple[TorchPredType, TorchFeatureType]:
        return model_outputs.forward(model, input)

    def _compute_preds_and_losses(
        self, model: nn.Module, optimizer: Optimizer, input: TorchInputType, target: TorchTargetType
    ) -> tuple[TrainingLosses, TorchPredType]:
        optimizer.zero_grad()
        with self._amp():
            preds, features = self.predict_with_model(model, input)
            return self.compute_training_loss(preds, features, self.transform_target(target)), preds

    def _apply_backwards_on_losses_and_take_step(
        self, model: nn.Module, optimizer: Optimizer, losses: TrainingLosses
    ) -> TrainingLosses:
        losses.backward["backward"].backward()
        self._transform_gradients_with_model(model, losses)
        optimizer.step()
        return losses

    def _train_step_with_model_and_optimizer(
        self, model: nn.Module, optimizer: Optimizer, input: TorchInputType, target: TorchTargetType
    ) -> tuple[TrainingLosses, TorchPredType]:
        losses, preds = self._compute_preds_and_losses(model, optimizer, input, target)
        return self._apply_backwards_on_losses_and_take_step(model, optimizer, losses), preds

    def _val_step_with_model(
        self, model: nn.Module, input: TorchInputType, target: TorchTargetType
    ) -> tuple[EvaluationLosses, TorchPredType]:
        with torch.no_grad(), self._amp():
            preds, features = self.predict_with_model(model, input)
            return self.compute_evaluation_loss(preds, features, self.transform_target(target)), preds

    def _transform_gradients_with_model(self, model: nn.Module, losses: TrainingLosses) -> None:
        pass

    # ---- the classic entry points: the helpers applied to (self.model, optimizers["global"]) -----------------
    def predict(self, input: TorchInputType) -> tuple[TorchPredType, TorchFeatureType]:
        return self.predict_with_model(self.model, input)

    def train_step(self, input: TorchInputType, target: TorchTargetType) -> tuple[TrainingLosses, TorchPredType]:
        return self._train_step_with_model_and_optimizer(self.model, self.optimizers["global"], input, target)

    def val_step(self, input: TorchInputType, target: TorchTargetType) -> tuple[EvaluationLosses, TorchPredType]:
        return self._val_step_with_model(self.model, input, target)

    def transform_gradients(self, losses: TrainingLosses) -> None:
        self._transform_gradients_with_model(self.model, losses)
