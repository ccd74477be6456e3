"""APFL client (Deng et al. 2020) over an ``ApflModule`` (parity: ``fl4health/clients/apfl_client.py:18-156``):
per step a global-model update, then a personal (alpha-mixed) update of the local model; alpha adapts on the first
step of each round; only ``global_model.*`` is exchanged."""

from __future__ import annotations

import torch
from torch.optim import Optimizer

from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.common.typing import Config
from fl4health_b200.model_bases.apfl_base import ApflModule
from fl4health_b200.parameter_exchange.layer_exchanger import FixedLayerExchanger
from fl4health_b200.utils.losses import TrainingLosses
from fl4health_b200.utils.typing import TorchFeatureType, TorchInputType, TorchPredType, TorchTargetType

_BRANCHES = ("global", "local")


class ApflClient(BasicClient):
    """Constructor arguments are ``BasicClient``'s; ``get_model`` must return an ``ApflModule`` and ``get_optimizer`` one
    optimizer per branch."""

    model: ApflModule
    learning_rate: float

    # ---------------------------------------------------------------------------------------------- wiring
    def get_optimizer(self, config: Config) -> dict[str, Optimizer]:
        raise NotImplementedError("Return {'local': opt(model.local_model.parameters()), 'global': opt(model.global_model.parameters())}")

    def set_optimizer(self, config: Config) -> None:
        per_branch = self.get_optimizer(config)
        assert isinstance(per_branch, dict) and set(per_branch) == set(_BRANCHES)
        self.optimizers = per_branch

    def get_parameter_exchanger(self, config: Config) -> FixedLayerExchanger:
        return FixedLayerExchanger(self.model.layers_to_exchange())

    # ---------------------------------------------------------------------------------------------- alpha
    def is_start_of_local_training(self, step: int) -> bool:
        return step == 0

    def update_after_step(self, step: int, current_round: int | None = None) -> None:
        if self.model.adaptive_alpha and self.is_start_of_local_training(step):
            self.model.update_alpha()  # consumes the gradients the first step of the round left behind

    def _graph_variant(self) -> object:
        # alpha is a Python float baked into captured kernels: one graph per value (it changes once per round at most)
        return round(float(self.model.alpha), 6)

    # ---------------------------------------------------------------------------------------------- step
    def compute_loss_and_additional_losses(
        self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType
    ) -> tuple[torch.Tensor, dict[str, torch.Tensor]]:
        return self.criterion(preds["personal"], target), {branch: self.criterion(preds[branch], target) for branch in _BRANCHES}

    def _global_branch_step(self, batch: torch.Tensor, target: TorchTargetType) -> None:
        optimizer = self.optimizers["global"]
        optimizer.zero_grad()
        with self._amp():
            objective = self.criterion(self.model.global_forward(batch), target)
        objective.backward()
        optimizer.step()

    def train_step(self, input: TorchInputType, target: TorchTargetType) -> tuple[TrainingLosses, TorchPredType]:
        assert isinstance(input, torch.Tensor)
        self._global_branch_step(input, target)
        # personal update: the alpha-mixed prediction is differentiated w.r.t. the LOCAL branch only
        for branch in _BRANCHES:
            self.optimizers[branch].zero_grad()
        with self._amp():
            preds, features = self.predict(input)
            losses = self.compute_training_loss(preds, features, self.transform_target(target))
        losses.backward["backward"].backward()
        self.optimizers["local"].step()
        return losses, preds
