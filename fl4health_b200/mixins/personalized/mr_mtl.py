"""MR-MTL as a mixin (parity: ``fl4health/mixins/personalized/mr_mtl.py:27-196``): the client's model is personal and
never overwritten after round 1; the server aggregate lands in ``initial_global_model`` and only anchors the penalty."""

from __future__ import annotations

from logging import INFO
from typing import Any

from torch import nn

from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config, NDArrays, Scalar
from fl4health_b200.mixins.adaptive_drift_constrained import AdaptiveDriftConstrainedMixin
from fl4health_b200.utils.losses import TrainingLosses
from fl4health_b200.utils.typing import TorchFeatureType, TorchPredType, TorchTargetType
from fl4health_b200.mixins.core_protocols import MrMtlPersonalizedProtocol  # noqa: F401  (import-path parity)


class MrMtlPersonalizedMixin(AdaptiveDriftConstrainedMixin):
    anchor_from_received_model = False

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        self.initial_global_model: nn.Module | None = None
        super().__init__(*args, **kwargs)

    def get_global_model(self, config: Config) -> nn.Module:
        return self.get_model(config)  # type: ignore[attr-defined]

    def setup_client(self, config: Config) -> None:
        if self.initial_global_model is None:
            self.initial_global_model = self._place_model(self.get_global_model(config), with_grad=False)  # type: ignore[attr-defined]
        super().setup_client(config)  # type: ignore[misc]

    def get_optimizer(self, config: Config) -> Any:
        """Hook kept from the reference (``mr_mtl.py:94-109``): a last chance to create the anchor model when a wrapped
        client builds its optimizer before ``setup_client`` ran; returns whatever the wrapped client returns."""
        if self.initial_global_model is None:
            self.initial_global_model = self._place_model(self.get_global_model(config), with_grad=False)  # type: ignore[attr-defined]
            log(INFO, f"initial_global_model set: {type(self.initial_global_model).__name__} within `get_optimizer`")
        return super().get_optimizer(config=config)  # type: ignore[misc]

    def set_parameters(self, parameters: NDArrays, config: Config, fitting_round: bool) -> None:
        assert self.initial_global_model is not None and self.parameter_exchanger is not None  # type: ignore[attr-defined]
        server_model_state, self.drift_penalty_weight = self.parameter_exchanger.unpack_parameters(parameters)  # type: ignore[attr-defined]
        log(INFO, f"Lambda weight received from the server: {self.drift_penalty_weight}")
        self.parameter_exchanger.pull_parameters(server_model_state, self.initial_global_model, config)  # type: ignore[attr-defined]

    def update_before_train(self, current_server_round: int) -> None:
        assert self.initial_global_model is not None
        for param in self.initial_global_model.parameters():
            param.requires_grad = False
        self.initial_global_model.eval()
        self.drift_penalty_tensors = self.snapshot_drift_anchor(source_model=self.initial_global_model)
        super().update_before_train(current_server_round)

    def compute_training_loss(
        self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType
    ) -> TrainingLosses:
        assert self.initial_global_model is not None and not self.initial_global_model.training and self.model.training  # type: ignore[attr-defined]
        return super().compute_training_loss(preds, features, target)  # type: ignore[misc]

    def validate(self, include_losses_in_metrics: bool = False) -> tuple[float, dict[str, Scalar]]:
        assert self.initial_global_model is not None and not self.initial_global_model.training
        return super().validate(include_losses_in_metrics=include_losses_in_metrics)  # type: ignore[misc]
