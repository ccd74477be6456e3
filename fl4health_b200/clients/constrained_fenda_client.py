"""FENDA with auxiliary feature-space constraints: cosine similarity between local/global features, MOON-style
contrastive loss, or PerFCL losses, using frozen copies of last round's extractors
(parity: ``fl4health/clients/constrained_fenda_client.py:22-267``).  Frozen extractors run without autograd."""

from __future__ import annotations

from typing import Any

import torch

from fl4health_b200.clients.fenda_client import FendaClient
from fl4health_b200.common.typing import Config
from fl4health_b200.engine.snapshots import SnapshotBank
from fl4health_b200.losses.fenda_loss_config import ConstrainedFendaLossContainer
from fl4health_b200.model_bases.fenda_base import FendaModelWithFeatureState
from fl4health_b200.parameter_exchange.parameter_exchanger_base import ParameterExchanger
from fl4health_b200.utils.losses import EvaluationLosses
from fl4health_b200.utils.typing import TorchFeatureType, TorchInputType, TorchPredType, TorchTargetType


# frozen slot -> feature key; the local extractor of last round serves both the contrastive and the PerFCL term
_REFERENCES = {"old_local": "old_local_features", "old_global": "old_global_features", "initial_global": "initial_global_features"}


class ConstrainedFendaClient(FendaClient):
    def __init__(self, *args: Any, loss_container: ConstrainedFendaLossContainer | None = None, **kwargs: Any) -> None:
        """``loss_container`` selects the auxiliary terms; every other argument is ``BasicClient``'s."""
        super().__init__(*args, **kwargs)
        self.loss_container = loss_container or ConstrainedFendaLossContainer(None, None, None)
        self._frozen = SnapshotBank(**{slot: 1 for slot in _REFERENCES})

    old_local_module = property(lambda self: self._frozen.get("old_local"))
    old_global_module = property(lambda self: self._frozen.get("old_global"))
    initial_global_module = property(lambda self: self._frozen.get("initial_global"))

    def _graph_variant(self) -> object:
        return self._frozen.variant()

    def get_parameter_exchanger(self, config: Config) -> ParameterExchanger:
        assert isinstance(self.model, FendaModelWithFeatureState)
        return super().get_parameter_exchanger(config)

    def _wanted_references(self) -> list[str]:
        """Frozen slots the configured loss terms actually read."""
        terms = self.loss_container
        slots = ["old_local"] if (terms.has_contrastive_loss() or terms.has_perfcl_loss()) else []
        return slots + (["old_global", "initial_global"] if terms.has_perfcl_loss() else [])

    # ------------------------------------------------------------------------------------------ round boundaries
    def update_before_train(self, current_server_round: int) -> None:
        assert isinstance(self.model, FendaModelWithFeatureState)
        if self.loss_container.has_perfcl_loss():
            self._frozen.capture("initial_global", self.model.second_feature_extractor)
        super().update_before_train(current_server_round)

    def update_after_train(self, local_steps: int, loss_dict: dict[str, float], config: Config) -> None:
        assert isinstance(self.model, FendaModelWithFeatureState)
        if self.loss_container.has_contrastive_loss() or self.loss_container.has_perfcl_loss():
            self._frozen.capture("old_local", self.model.first_feature_extractor)
            self._frozen.capture("old_global", self.model.second_feature_extractor)
        super().update_after_train(local_steps, loss_dict, config)

    # ------------------------------------------------------------------------------------------ step
    def predict(self, input: TorchInputType) -> tuple[TorchPredType, TorchFeatureType]:
        assert isinstance(input, torch.Tensor) and isinstance(self.model, FendaModelWithFeatureState)
        preds, features = self.model(input)
        with torch.no_grad():
            for slot in self._wanted_references():
                frozen = self._frozen.get(slot)
                if frozen is not None:
                    reference = frozen(input)
                    features[_REFERENCES[slot]] = reference.reshape(len(reference), -1)
        return preds, features

    def compute_loss_and_additional_losses(
        self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType
    ) -> tuple[torch.Tensor, dict[str, torch.Tensor]]:
        terms = self.loss_container
        task = self.criterion(preds["prediction"], target)
        recorded: dict[str, torch.Tensor] = {"loss": task}
        extras: list[torch.Tensor] = []
        if terms.has_cosine_similarity_loss():
            recorded["cos_sim_loss"] = terms.compute_cosine_similarity_loss(features["local_features"], features["global_features"])
            extras.append(recorded["cos_sim_loss"])
        if terms.has_contrastive_loss() and "old_local_features" in features:
            recorded["contrastive_loss"] = terms.compute_contrastive_loss(
                features["local_features"], features["old_local_features"].unsqueeze(0), features["global_features"].unsqueeze(0))
            extras.append(recorded["contrastive_loss"])
        if terms.has_perfcl_loss() and set(_REFERENCES.values()) <= features.keys():
            pull_global, push_local = terms.compute_perfcl_loss(
                features["local_features"], features["old_local_features"], features["global_features"],
                features["old_global_features"], features["initial_global_features"])
            recorded["global_feature_contrastive_loss"], recorded["local_feature_contrastive_loss"] = pull_global, push_local
            extras += [pull_global, push_local]
        total = task.clone()
        for extra in extras:
            total = total + extra
        recorded["total_loss"] = total
        return total, recorded

    def compute_evaluation_loss(self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType) -> EvaluationLosses:
        recorded = self.compute_loss_and_additional_losses(preds, features, target)[1]
        return EvaluationLosses(checkpoint=recorded["loss"], additional_losses=recorded)
