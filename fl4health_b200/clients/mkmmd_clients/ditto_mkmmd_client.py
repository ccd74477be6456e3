"""Ditto with an MK-MMD penalty between personal-model features and round-start global-model features
(parity: ``fl4health/clients/mkmmd_clients/ditto_mkmmd_client.py:21-359``)."""

from __future__ import annotations

from collections.abc import Sequence
from pathlib import Path

import torch

from fl4health_b200.checkpointing.client_module import ClientCheckpointAndStateModule
from fl4health_b200.common.typing import Config
from fl4health_b200.engine.options import EngineOptions
from fl4health_b200.metrics.base_metrics import Metric
from fl4health_b200.reporting.base_reporter import BaseReporter
from fl4health_b200.utils.client import clone_and_freeze_model
from fl4health_b200.utils.losses import LossMeterType, TrainingLosses
from fl4health_b200.utils.typing import TorchFeatureType, TorchInputType, TorchPredType, TorchTargetType
from fl4health_b200.clients._mmd_feature_alignment import MkMmdMixin
from fl4health_b200.clients.ditto_client import DittoClient


class DittoMkMmdClient(MkMmdMixin, DittoClient):
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
        mkmmd_loss_weight: float = 10.0,
        feature_extraction_layers: Sequence[str] | None = None,
        feature_l2_norm_weight: float = 0.0,
        beta_global_update_interval: int = 20,
        num_accumulating_batches: int | None = None,
        engine_options: EngineOptions | None = None,
    ) -> None:
        """``beta_global_update_interval``: -1 = re-optimise the kernel weights on every batch, 0 = never, n > 0 = every
        n steps from features accumulated over ``num_accumulating_batches`` training batches."""
        DittoClient.__init__(
            self, data_path=data_path, metrics=metrics, device=device, loss_meter_type=loss_meter_type,
            checkpoint_and_state_module=checkpoint_and_state_module, reporters=reporters, progress_bar=progress_bar,
            client_name=client_name, engine_options=engine_options,
        )
        self._init_mkmmd(mkmmd_loss_weight, feature_extraction_layers, feature_l2_norm_weight, beta_global_update_interval,
                         num_accumulating_batches)

    def setup_client(self, config: Config) -> None:
        super().setup_client(config)
        self._attach_local_hooks()

    def update_before_train(self, current_server_round: int) -> None:
        super().update_before_train(current_server_round)
        # frozen copy of the GLOBAL model as received this round: the feature anchor (the live global model keeps training)
        self.initial_global_model = clone_and_freeze_model(self.global_model)
        self._attach_anchor_hooks()

    def predict(self, input: TorchInputType) -> tuple[TorchPredType, TorchFeatureType]:
        preds, _ = super().predict(input)  # global + personal forward (hooks fire on the personal model)
        return preds, self._collect_features(input)

    def compute_loss_and_additional_losses(
        self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType
    ) -> tuple[torch.Tensor, dict[str, torch.Tensor]]:
        loss, additional = super().compute_loss_and_additional_losses(preds, features, target)
        additional.update(self._mmd_terms(features))
        additional.update(self._feature_norm_term(features))
        return loss, additional

    def compute_training_loss(
        self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType
    ) -> TrainingLosses:
        assert self.global_model.training and self.model.training
        loss, additional = self.compute_loss_and_additional_losses(preds, features, target)
        additional["loss_for_adaptation"] = additional["local_loss"].clone()
        penalty = self.compute_penalty_loss()
        additional["penalty_loss"] = penalty.clone()
        total = loss + penalty
        for key in ("mkmmd_loss_total", "feature_l2_norm_loss"):
            if key in additional:
                total = total + additional[key]
        additional["total_loss"] = total.clone()
        return TrainingLosses(backward=total, additional_losses=additional)
