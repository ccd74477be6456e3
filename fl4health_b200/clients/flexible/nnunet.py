"""nnU-Net on the flexible client API (parity: ``fl4health/clients/flexible/nnunet.py:85``): the model / optimizer
parameterised helpers carry nnU-Net's specifics (autocast forward, deep-supervision dicts, GradScaler step, gradient
clipping), so the personalisation mixins (``make_it_personal(FlexibleNnunetClient, DITTO)``) can drive a second model
through exactly the same code path."""

from __future__ import annotations

import torch
from torch import nn
from torch.optim import Optimizer

from fl4health_b200.clients.flexible.base import FlexibleClient
from fl4health_b200.clients.nnunet_client import NnunetClient
from fl4health_b200.utils.losses import TrainingLosses
from fl4health_b200.utils.nnunet_utils import NNUNET_N_SPATIAL_DIMS, convert_deep_supervision_list_to_dict
from fl4health_b200.utils.typing import TorchFeatureType, TorchInputType, TorchPredType, TorchTargetType


class FlexibleNnunetClient(NnunetClient, FlexibleClient):
    _dynamically_created = True  # the legacy hooks below are re-routed to the helpers on purpose

    def predict_with_model(self, model: nn.Module, input: TorchInputType) -> tuple[TorchPredType, TorchFeatureType]:
        if not isinstance(input, torch.Tensor):
            raise TypeError('"input" must be of type torch.Tensor for nnUNetClient')
        with torch.autocast(self.device.type, enabled=self.device.type == "cuda"):
            output = model(input)
        if isinstance(output, torch.Tensor):
            return {"prediction": output}, {}
        if isinstance(output, (list, tuple)):
            return convert_deep_supervision_list_to_dict(output, NNUNET_N_SPATIAL_DIMS[self.nnunet_config]), {}
        raise TypeError("Was expecting nnunet model output to be either a torch.Tensor or a list/tuple of torch.Tensors")

    def predict(self, input: TorchInputType) -> tuple[TorchPredType, TorchFeatureType]:
        return self.predict_with_model(self.model, input)

    def _transform_gradients_with_model(self, model: nn.Module, losses: TrainingLosses) -> None:
        nn.utils.clip_grad_norm_(model.parameters(), self.max_grad_norm)

    def transform_gradients(self, losses: TrainingLosses) -> None:
        self._transform_gradients_with_model(self.model, losses)

    def _apply_backwards_on_losses_and_take_step(self, model: nn.Module, optimizer: Optimizer, losses: TrainingLosses) -> TrainingLosses:
        if self.device.type != "cuda":
            return FlexibleClient._apply_backwards_on_losses_and_take_step(self, model, optimizer, losses)
        self.grad_scaler.scale(losses.backward["backward"]).backward()
        self.grad_scaler.unscale_(optimizer)
        self._transform_gradients_with_model(model, losses)
        self.grad_scaler.step(optimizer)
        self.grad_scaler.update()
        return losses

    def train_step(self, input: TorchInputType, target: TorchTargetType) -> tuple[TrainingLosses, TorchPredType]:
        return self._train_step_with_model_and_optimizer(self.model, self.optimizers["global"], input, target)

    def val_step(self, input: TorchInputType, target: TorchTargetType):  # noqa: ANN201
        return self._val_step_with_model(self.model, input, target)
