"""MR-MTL with a learned deep-kernel MMD feature penalty (parity:
``fl4health/clients/deep_mmd_clients/mr_mtl_deep_mmd_client.py``)."""

from __future__ import annotations

from collections.abc import Sequence
from pathlib import Path

import torch

from fl4health_b200.checkpointing.client_module import ClientCheckpointAndStateModule
from fl4health_b200.common.typing import Config, Scalar
from fl4health_b200.engine.options import EngineOptions
from fl4health_b200.metrics.base_metrics import Metric
from fl4health_b200.reporting.base_reporter import BaseReporter
from fl4health_b200.utils.losses import EvaluationLosses, LossMeterType, TrainingLosses
from fl4health_b200.utils.typing import TorchFeatureType, TorchInputType, TorchPredType, TorchTargetType
from fl4health_b200.clients._mmd_feature_alignment import DeepMmdMixin
from fl4health_b200.clients.mr_mtl_client import MrMtlClient


class MrMtlDeepMmdClient(DeepMmdMixin, MrMtlClient):
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
        deep_mmd_loss_weight: float = 10.0,
        feature_extraction_layers_with_size: dict[str, int] | None = None,
        mmd_kernel_train_interval: int = 20,
        num_accumulating_batches: int | None = None,
        engine_options: EngineOptions | None = None,
    ) -> None:
        MrMtlClient.__init__(
            self, data_path=data_path, metrics=metrics, device=device, loss_meter_type=loss_meter_type,
            checkpoint_and_state_module=checkpoint_and_state_module, reporters=reporters, progress_bar=progress_bar,
            client_name=client_name, engine_options=engine_options,
        )
        self._init_deep_mmd(deep_mmd_loss_weight, feature_extraction_layers_with_size, mmd_kernel_train_interval,
                            num_accumulating_batches)

    def setup_client(self, config: Config) -> None:
        super().setup_client(config)
        self._attach_local_hooks()

    def update_before_train(self, current_server_round: int) -> None:
        super().update_before_train(current_server_round)
        self._attach_anchor_hooks()
        self._set_kernel_training(self.mmd_kernel_train_interval == -1)

    def predict(self, input: TorchInputType) -> tuple[TorchPredType, TorchFeatureType]:
        preds, _ = super().predict(input)
        return preds, self._collect_features(input)

    def validate(self, include_losses_in_metrics: bool = False) -> tuple[float, dict[str, Scalar]]:
        self._set_kernel_training(False)
        return super().validate(include_losses_in_metrics)

    def compute_loss_and_additional_losses(
        self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType
    ) -> tuple[torch.Tensor, dict[str, torch.Tensor]]:
        loss, additional = super().compute_loss_and_additional_losses(preds, features, target)
        additional = additional if additional is not None else {"loss": loss}
        additional.update(self._mmd_terms(features))
        return loss, additional

    def compute_training_loss(
        self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType
    ) -> TrainingLosses:
        assert self.model.training
        loss, additional = self.compute_loss_and_additional_losses(preds, features, target)
        additional["loss"] = loss.clone()
        additional["loss_for_adaptation"] = loss.clone()
        penalty = self.compute_penalty_loss()
        additional["penalty_loss"] = penalty.clone()
        total = loss + penalty + additional.get("deep_mmd_loss_total", 0.0)
        additional["total_loss"] = total.clone()
        return TrainingLosses(backward=total, additional_losses=additional)

    def compute_evaluation_loss(
        self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType
    ) -> EvaluationLosses:
        for loss in self.deep_mmd_losses.values():
            assert not loss.training
        return super().compute_evaluation_loss(preds, features, target)
