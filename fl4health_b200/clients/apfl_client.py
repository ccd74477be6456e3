"""APFL client (Deng et al. 2020) over an ``ApflModule`` (parity: ``fl4health/clients/apfl_client.py:18-156``):
per step a global-model update, then a personal (alpha-mixed) update of the local model; alpha adapts on the first
step of each round; only ``global_model.*`` is exchanged."""

from __future__ import annotations

from collections.abc import Sequence
from pathlib import Path

import torch
from torch.optim import Optimizer

from fl4health_b200.checkpointing.client_module import ClientCheckpointAndStateModule
from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.common.typing import Config
from fl4health_b200.engine.options import EngineOptions
from fl4health_b200.metrics.base_metrics import Metric
from fl4health_b200.model_bases.apfl_base import ApflModule
from fl4health_b200.parameter_exchange.layer_exchanger import FixedLayerExchanger
from fl4health_b200.reporting.base_reporter import BaseReporter
from fl4health_b200.utils.losses import LossMeterType, TrainingLosses
from fl4health_b200.utils.typing import TorchFeatureType, TorchInputType, TorchPredType, TorchTargetType


class ApflClient(BasicClient):
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
        self.model: ApflModule
        self.learning_rate: float

    def is_start_of_local_training(self, step: int) -> bool:
        return step == 0

    def update_after_step(self, step: int, current_round: int | None = None) -> None:
        if self.is_start_of_local_training(step) and self.model.adaptive_alpha:
            self.model.update_alpha()  # reads the gradients left by the step that just ran

    def _graph_variant(self) -> object:
        # alpha is a Python float baked into captured kernels: re-capture when it changes (once per round at most)
        return round(float(self.model.alpha), 6)

    def train_step(self, input: TorchInputType, target: TorchTargetType) -> tuple[TrainingLosses, TorchPredType]:
        assert isinstance(input, torch.Tensor)
        # (1) global model step
        self.optimizers["global"].zero_grad()
        with self._amp():
            global_loss = self.criterion(self.model.global_forward(input), target)
        global_loss.backward()
        self.optimizers["global"].step()
        # (2) personal step on the local model
        self.optimizers["global"].zero_grad()
        self.optimizers["local"].zero_grad()
        with self._amp():
            preds, features = self.predict(input)
            target = self.transform_target(target)
            losses = self.compute_training_loss(preds, features, target)
        losses.backward["backward"].backward()
        self.optimizers["local"].step()
        return losses, preds

    def get_parameter_exchanger(self, config: Config) -> FixedLayerExchanger:
        return FixedLayerExchanger(self.model.layers_to_exchange())

    def compute_loss_and_additional_losses(
        self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType
    ) -> tuple[torch.Tensor, dict[str, torch.Tensor]]:
        personal_loss = self.criterion(preds["personal"], target)
        return personal_loss, {"global": self.criterion(preds["global"], target), "local": self.criterion(preds["local"], target)}

    def set_optimizer(self, config: Config) -> None:
        optimizers = self.get_optimizer(config)
        assert isinstance(optimizers, dict) and set(optimizers.keys()) == {"global", "local"}
        self.optimizers = optimizers

    def get_optimizer(self, config: Config) -> dict[str, Optimizer]:
        raise NotImplementedError("Return {'local': opt(model.local_model.parameters()), 'global': opt(model.global_model.parameters())}")
