"""MR-MTL (mean-regularised multi-task learning): only a personal model is trained; the server aggregate is kept in
``initial_global_model`` purely as the drift anchor (parity: ``fl4health/clients/mr_mtl_client.py:18-167``)."""

from __future__ import annotations

from collections.abc import Sequence
from logging import INFO
from pathlib import Path

import torch
from torch import nn

from fl4health_b200.checkpointing.client_module import ClientCheckpointAndStateModule
from fl4health_b200.clients.adaptive_drift_constraint_client import AdaptiveDriftConstraintClient
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config, NDArrays, Scalar
from fl4health_b200.engine.options import EngineOptions
from fl4health_b200.metrics.base_metrics import Metric
from fl4health_b200.reporting.base_reporter import BaseReporter
from fl4health_b200.utils.losses import LossMeterType, TrainingLosses
from fl4health_b200.utils.typing import TorchFeatureType, TorchPredType, TorchTargetType


class MrMtlClient(AdaptiveDriftConstraintClient):
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
        self.initial_global_model: nn.Module
        self.initial_global_tensors: list[torch.Tensor]

    def setup_client(self, config: Config) -> None:
        self.initial_global_model = self._place_model(self.get_model(config), with_grad=False)
        super().setup_client(config)

    def set_parameters(self, parameters: NDArrays, config: Config, fitting_round: bool) -> None:
        """The aggregate never overwrites the personal model: it lands in ``initial_global_model``."""
        assert self.initial_global_model is not None and self.parameter_exchanger is not None
        server_model_state, self.drift_penalty_weight = self.parameter_exchanger.unpack_parameters(parameters)
        log(INFO, f"Lambda weight received from the server: {self.drift_penalty_weight}")
        self.parameter_exchanger.pull_parameters(server_model_state, self.initial_global_model, config)

    def update_before_train(self, current_server_round: int) -> None:
        for param in self.initial_global_model.parameters():
            param.requires_grad = False
        self.initial_global_model.eval()
        self.drift_penalty_tensors = self.snapshot_drift_anchor(
            source_model=self.initial_global_model, constrained_model=self.model
        )
        return super().update_before_train(current_server_round)

    def compute_training_loss(
        self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType
    ) -> TrainingLosses:
        assert not self.initial_global_model.training and self.model.training
        return super().compute_training_loss(preds, features, target)

    def validate(self, include_losses_in_metrics: bool = False) -> tuple[float, dict[str, Scalar]]:
        assert not self.initial_global_model.training
        return super().validate(include_losses_in_metrics=include_losses_in_metrics)
