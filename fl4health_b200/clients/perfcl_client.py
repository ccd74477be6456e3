"""PerFCL client (parity: ``fl4health/clients/perfcl_client.py:20-278``): FENDA-style model, two contrastive terms
computed against frozen copies of last round's local/global extractors and the round-start global extractor."""

from __future__ import annotations

from typing import Any

import torch

from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.common.typing import Config
from fl4health_b200.engine.snapshots import SnapshotBank
from fl4health_b200.losses.perfcl_loss import PerFclLoss
from fl4health_b200.model_bases.perfcl_base import PerFclModel
from fl4health_b200.parameter_exchange.layer_exchanger import FixedLayerExchanger
from fl4health_b200.parameter_exchange.parameter_exchanger_base import ParameterExchanger
from fl4health_b200.utils.losses import EvaluationLosses
from fl4health_b200.utils.typing import TorchFeatureType, TorchInputType, TorchPredType, TorchTargetType

# frozen slot -> feature key it contributes to the step
_REFERENCES = {"old_local": "old_local_features", "old_global": "old_global_features", "initial_global": "initial_global_features"}


class PerFclClient(BasicClient):
    def __init__(
        self, *args: Any, global_feature_loss_temperature: float = 0.5, local_feature_loss_temperature: float = 0.5,
        global_feature_contrastive_loss_weight: float = 1.0, local_feature_contrastive_loss_weight: float = 1.0, **kwargs: Any,
    ) -> None:
        """Arguments other than the four PerFCL knobs are ``BasicClient``'s."""
        super().__init__(*args, **kwargs)
        self.global_feature_contrastive_loss_weight = global_feature_contrastive_loss_weight
        self.local_feature_contrastive_loss_weight = local_feature_contrastive_loss_weight
        self.perfcl_loss_function = PerFclLoss(self.device, global_feature_loss_temperature, local_feature_loss_temperature)
        self._frozen = SnapshotBank(**{slot: 1 for slot in _REFERENCES})

    old_local_module = property(lambda self: self._frozen.get("old_local"))
    old_global_module = property(lambda self: self._frozen.get("old_global"))
    initial_global_module = property(lambda self: self._frozen.get("initial_global"))

    def _graph_variant(self) -> object:
        return self._frozen.variant()

    def _all_contrastive_loss_modules_defined(self) -> bool:
        return self._frozen.filled()

    def get_parameter_exchanger(self, config: Config) -> ParameterExchanger:
        assert isinstance(self.model, PerFclModel)
        return FixedLayerExchanger(self.model.layers_to_exchange())

    # ---------------------------------------------------------------------------------------------- round boundaries
    def update_before_train(self, current_server_round: int) -> None:
        assert isinstance(self.model, PerFclModel)
        self._frozen.capture("initial_global", self.model.second_feature_extractor)
        super().update_before_train(current_server_round)

    def update_after_train(self, local_steps: int, loss_dict: dict[str, float], config: Config) -> None:
        assert isinstance(self.model, PerFclModel)
        self._frozen.capture("old_local", self.model.first_feature_extractor)
        self._frozen.capture("old_global", self.model.second_feature_extractor)
        super().update_after_train(local_steps, loss_dict, config)

    # ---------------------------------------------------------------------------------------------- step
    def predict(self, input: TorchInputType) -> tuple[TorchPredType, TorchFeatureType]:
        assert isinstance(input, torch.Tensor)
        preds, features = self.model(input)
        if self._frozen.filled():
            with torch.no_grad():
                for slot, key in _REFERENCES.items():
                    reference = self._frozen.get(slot)(input)  # type: ignore[misc]
                    features[key] = reference.reshape(len(reference), -1)
        return preds, features

    def compute_loss_and_additional_losses(
        self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType
    ) -> tuple[torch.Tensor, dict[str, torch.Tensor]]:
        task = self.criterion(preds["prediction"], target)
        if not (self._frozen.filled() and "old_local_features" in features):
            return task, {"loss": task}
        pull_global, push_local = self.perfcl_loss_function(
            features["local_features"], features["old_local_features"], features["global_features"],
            features["old_global_features"], features["initial_global_features"])
        total = (task + self.global_feature_contrastive_loss_weight * pull_global
                 + self.local_feature_contrastive_loss_weight * push_local)
        return total, {"loss": task, "global_feature_contrastive_loss": pull_global,
                       "local_feature_contrastive_loss": push_local, "total_loss": total}

    def compute_evaluation_loss(self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType) -> EvaluationLosses:
        recorded = self.compute_loss_and_additional_losses(preds, features, target)[1]
        return EvaluationLosses(checkpoint=recorded["loss"], additional_losses=recorded)
