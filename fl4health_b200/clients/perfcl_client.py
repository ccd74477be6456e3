"""PerFCL client (parity: ``fl4health/clients/perfcl_client.py:20-278``): FENDA-style model, two contrastive terms
computed against frozen copies of last round's local/global extractors and the round-start global extractor."""

from __future__ import annotations

from collections.abc import Sequence
from pathlib import Path
from typing import Any

import torch

from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.common.typing import Config
from fl4health_b200.losses.perfcl_loss import PerFclLoss
from fl4health_b200.metrics.base_metrics import Metric
from fl4health_b200.model_bases.perfcl_base import PerFclModel
from fl4health_b200.parameter_exchange.layer_exchanger import FixedLayerExchanger
from fl4health_b200.parameter_exchange.parameter_exchanger_base import ParameterExchanger
from fl4health_b200.utils.client import clone_and_freeze_model
from fl4health_b200.utils.losses import EvaluationLosses, LossMeterType
from fl4health_b200.utils.typing import TorchFeatureType, TorchInputType, TorchPredType, TorchTargetType


class PerFclClient(BasicClient):
    def __init__(
        self,
        data_path: Path,
        metrics: Sequence[Metric],
        device: torch.device,
        loss_meter_type: LossMeterType = LossMeterType.AVERAGE,
        checkpoint_and_state_module: Any = None,
        reporters: Any = None,
        progress_bar: bool = False,
        client_name: str | None = None,
        global_feature_loss_temperature: float = 0.5,
        local_feature_loss_temperature: float = 0.5,
        global_feature_contrastive_loss_weight: float = 1.0,
        local_feature_contrastive_loss_weight: float = 1.0,
        engine_options: Any = None,
    ) -> None:
        super().__init__(data_path=data_path, metrics=metrics, device=device, loss_meter_type=loss_meter_type,
                         checkpoint_and_state_module=checkpoint_and_state_module, reporters=reporters,
                         progress_bar=progress_bar, client_name=client_name, engine_options=engine_options)
        self.global_feature_contrastive_loss_weight = global_feature_contrastive_loss_weight
        self.local_feature_contrastive_loss_weight = local_feature_contrastive_loss_weight
        self.perfcl_loss_function = PerFclLoss(self.device, global_feature_loss_temperature, local_feature_loss_temperature)
        self.old_local_module: torch.nn.Module | None = None
        self.old_global_module: torch.nn.Module | None = None
        self.initial_global_module: torch.nn.Module | None = None

    def _graph_variant(self) -> object:
        return tuple(id(m) for m in (self.old_local_module, self.old_global_module, self.initial_global_module))

    def get_parameter_exchanger(self, config: Config) -> ParameterExchanger:
        assert isinstance(self.model, PerFclModel)
        return FixedLayerExchanger(self.model.layers_to_exchange())

    def _flatten(self, features: torch.Tensor) -> torch.Tensor:
        return features.reshape(len(features), -1)

    def _all_contrastive_loss_modules_defined(self) -> bool:
        return None not in (self.old_local_module, self.old_global_module, self.initial_global_module)

    def predict(self, input: TorchInputType) -> tuple[TorchPredType, TorchFeatureType]:
        assert isinstance(input, torch.Tensor)
        preds, features = self.model(input)
        if self._all_contrastive_loss_modules_defined():
            with torch.no_grad():
                features["old_local_features"] = self._flatten(self.old_local_module(input))  # type: ignore[misc]
                features["old_global_features"] = self._flatten(self.old_global_module(input))  # type: ignore[misc]
                features["initial_global_features"] = self._flatten(self.initial_global_module(input))  # type: ignore[misc]
        return preds, features

    def update_after_train(self, local_steps: int, loss_dict: dict[str, float], config: Config) -> None:
        assert isinstance(self.model, PerFclModel)
        self.old_local_module = clone_and_freeze_model(self.model.first_feature_extractor)
        self.old_global_module = clone_and_freeze_model(self.model.second_feature_extractor)
        super().update_after_train(local_steps, loss_dict, config)

    def update_before_train(self, current_server_round: int) -> None:
        assert isinstance(self.model, PerFclModel)
        self.initial_global_module = clone_and_freeze_model(self.model.second_feature_extractor)
        super().update_before_train(current_server_round)

    def compute_loss_and_additional_losses(
        self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType
    ) -> tuple[torch.Tensor, dict[str, torch.Tensor]]:
        loss = self.criterion(preds["prediction"], target)
        if not self._all_contrastive_loss_modules_defined() or "old_local_features" not in features:
            return loss, {"loss": loss}
        g, l = self.perfcl_loss_function(features["local_features"], features["old_local_features"],
                                         features["global_features"], features["old_global_features"],
                                         features["initial_global_features"])
        total = loss + self.global_feature_contrastive_loss_weight * g + self.local_feature_contrastive_loss_weight * l
        return total, {"loss": loss, "global_feature_contrastive_loss": g, "local_feature_contrastive_loss": l,
                       "total_loss": total}

    def compute_evaluation_loss(self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType) -> EvaluationLosses:
        _, additional = self.compute_loss_and_additional_losses(preds, features, target)
        return EvaluationLosses(checkpoint=additional["loss"], additional_losses=additional)
