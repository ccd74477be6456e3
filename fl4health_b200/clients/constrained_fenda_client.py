"""FENDA with auxiliary feature-space constraints: cosine similarity between local/global features, MOON-style
contrastive loss, or PerFCL losses, using frozen copies of last round's extractors
(parity: ``fl4health/clients/constrained_fenda_client.py:22-267``).  Frozen extractors run without autograd."""

from __future__ import annotations

from collections.abc import Sequence
from pathlib import Path
from typing import Any

import torch

from fl4health_b200.clients.fenda_client import FendaClient
from fl4health_b200.common.typing import Config
from fl4health_b200.losses.fenda_loss_config import ConstrainedFendaLossContainer
from fl4health_b200.metrics.base_metrics import Metric
from fl4health_b200.model_bases.fenda_base import FendaModelWithFeatureState
from fl4health_b200.parameter_exchange.parameter_exchanger_base import ParameterExchanger
from fl4health_b200.utils.client import clone_and_freeze_model
from fl4health_b200.utils.losses import EvaluationLosses, LossMeterType
from fl4health_b200.utils.typing import TorchFeatureType, TorchInputType, TorchPredType, TorchTargetType


class ConstrainedFendaClient(FendaClient):
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
        loss_container: ConstrainedFendaLossContainer | None = None,
        engine_options: Any = None,
    ) -> None:
        super().__init__(data_path=data_path, metrics=metrics, device=device, loss_meter_type=loss_meter_type,
                         checkpoint_and_state_module=checkpoint_and_state_module, reporters=reporters,
                         progress_bar=progress_bar, client_name=client_name, engine_options=engine_options)
        self.loss_container = loss_container if loss_container is not None else ConstrainedFendaLossContainer(None, None, None)
        self.old_local_module: torch.nn.Module | None = None
        self.old_global_module: torch.nn.Module | None = None
        self.initial_global_module: torch.nn.Module | None = None

    def _graph_variant(self) -> object:
        return tuple(id(m) for m in (self.old_local_module, self.old_global_module, self.initial_global_module))

    def get_parameter_exchanger(self, config: Config) -> ParameterExchanger:
        assert isinstance(self.model, FendaModelWithFeatureState)
        return super().get_parameter_exchanger(config)

    def _flatten(self, features: torch.Tensor) -> torch.Tensor:
        return features.reshape(len(features), -1)

    def _perfcl_keys_present(self, features: dict[str, torch.Tensor]) -> bool:
        return {"old_local_features", "old_global_features", "initial_global_features"} <= features.keys()

    def predict(self, input: TorchInputType) -> tuple[TorchPredType, TorchFeatureType]:
        assert isinstance(input, torch.Tensor) and isinstance(self.model, FendaModelWithFeatureState)
        preds, features = self.model(input)
        lc = self.loss_container
        with torch.no_grad():
            if (lc.has_contrastive_loss() or lc.has_perfcl_loss()) and self.old_local_module is not None:
                features["old_local_features"] = self._flatten(self.old_local_module(input))
            if lc.has_perfcl_loss():
                if self.old_global_module is not None:
                    features["old_global_features"] = self._flatten(self.old_global_module(input))
                if self.initial_global_module is not None:
                    features["initial_global_features"] = self._flatten(self.initial_global_module(input))
        return preds, features

    def update_after_train(self, local_steps: int, loss_dict: dict[str, float], config: Config) -> None:
        assert isinstance(self.model, FendaModelWithFeatureState)
        if self.loss_container.has_contrastive_loss() or self.loss_container.has_perfcl_loss():
            self.old_local_module = clone_and_freeze_model(self.model.first_feature_extractor)
            self.old_global_module = clone_and_freeze_model(self.model.second_feature_extractor)
        super().update_after_train(local_steps, loss_dict, config)

    def update_before_train(self, current_server_round: int) -> None:
        assert isinstance(self.model, FendaModelWithFeatureState)
        if self.loss_container.has_perfcl_loss():
            self.initial_global_module = clone_and_freeze_model(self.model.second_feature_extractor)
        super().update_before_train(current_server_round)

    def compute_loss_and_additional_losses(
        self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType
    ) -> tuple[torch.Tensor, dict[str, torch.Tensor]]:
        loss = self.criterion(preds["prediction"], target)
        total_loss = loss.clone()
        additional = {"loss": loss}
        lc = self.loss_container
        if lc.has_cosine_similarity_loss():
            cos = lc.compute_cosine_similarity_loss(features["local_features"], features["global_features"])
            total_loss = total_loss + cos
            additional["cos_sim_loss"] = cos
        if lc.has_contrastive_loss() and "old_local_features" in features:
            con = lc.compute_contrastive_loss(features["local_features"], features["old_local_features"].unsqueeze(0),
                                              features["global_features"].unsqueeze(0))
            total_loss = total_loss + con
            additional["contrastive_loss"] = con
        if lc.has_perfcl_loss() and self._perfcl_keys_present(features):
            g, l = lc.compute_perfcl_loss(features["local_features"], features["old_local_features"],
                                          features["global_features"], features["old_global_features"],
                                          features["initial_global_features"])
            total_loss = total_loss + g + l
            additional["global_feature_contrastive_loss"] = g
            additional["local_feature_contrastive_loss"] = l
        additional["total_loss"] = total_loss
        return total_loss, additional

    def compute_evaluation_loss(self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType) -> EvaluationLosses:
        _, additional = self.compute_loss_and_additional_losses(preds, features, target)
        return EvaluationLosses(checkpoint=additional["loss"], additional_losses=additional)
