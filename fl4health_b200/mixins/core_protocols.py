"""Structural types the mixins rely on (parity: ``fl4health/mixins/core_protocols.py:15-107``)."""

from __future__ import annotations

from typing import Any, Protocol, runtime_checkable

import torch
from torch import nn
from torch.optim import Optimizer

from fl4health_b200.common.typing import Config, NDArrays, Scalar
from fl4health_b200.utils.losses import EvaluationLosses, TrainingLosses
from fl4health_b200.utils.typing import TorchFeatureType, TorchInputType, TorchPredType, TorchTargetType


@runtime_checkable
class NumPyClientMinimalProtocol(Protocol):
    def get_parameters(self, config: dict[str, Scalar]) -> NDArrays: ...

    def fit(self, parameters: NDArrays, config: dict[str, Scalar]) -> tuple[NDArrays, int, dict[str, Scalar]]: ...

    def evaluate(self, parameters: NDArrays, config: dict[str, Scalar]) -> tuple[float, int, dict[str, Scalar]]: ...

    def set_parameters(self, parameters: NDArrays, config: Config, fitting_round: bool) -> None: ...

    def update_after_train(self, local_steps: int, loss_dict: dict[str, float], config: Config) -> None: ...


@runtime_checkable
class FlexibleClientProtocolPreSetup(NumPyClientMinimalProtocol, Protocol):
    device: torch.device
    initialized: bool

    def setup_client(self, config: Config) -> None: ...

    def get_model(self, config: Config) -> nn.Module: ...

    def get_data_loaders(self, config: Config) -> tuple[Any, ...]: ...

    def get_optimizer(self, config: Config) -> Optimizer | dict[str, Optimizer]: ...

    def get_criterion(self, config: Config) -> Any: ...

    def compute_loss_and_additional_losses(
        self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType
    ) -> tuple[torch.Tensor, dict[str, torch.Tensor] | None]: ...


@runtime_checkable
class FlexibleClientProtocol(FlexibleClientProtocolPreSetup, Protocol):
    model: nn.Module
    optimizers: dict[str, Optimizer]

    def initialize_all_model_weights(self, parameters: NDArrays, config: Config) -> None: ...

    def update_before_train(self, current_server_round: int) -> None: ...

    def _compute_preds_and_losses(
        self, model: nn.Module, optimizer: Optimizer, input: TorchInputType, target: TorchTargetType
    ) -> tuple[TrainingLosses, TorchPredType]: ...

    def _apply_backwards_on_losses_and_take_step(
        self, model: nn.Module, optimizer: Optimizer, losses: TrainingLosses
    ) -> TrainingLosses: ...

    def _train_step_with_model_and_optimizer(
        self, model: nn.Module, optimizer: Optimizer, input: TorchInputType, target: TorchTargetType
    ) -> tuple[TrainingLosses, TorchPredType]: ...

    def _val_step_with_model(
        self, model: nn.Module, input: TorchInputType, target: TorchTargetType
    ) -> tuple[EvaluationLosses, TorchPredType]: ...

    def predict_with_model(self, model: nn.Module, input: TorchInputType) -> tuple[TorchPredType, TorchFeatureType]: ...

    def transform_target(self, target: TorchTargetType) -> TorchTargetType: ...

    def _transform_gradients_with_model(self, model: nn.Module, losses: TrainingLosses) -> None: ...

    def transform_gradients(self, losses: TrainingLosses) -> None: ...

    def compute_training_loss(
        self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType
    ) -> TrainingLosses: ...

    def validate(self, include_losses_in_metrics: bool = False) -> tuple[float, dict[str, Scalar]]: ...

    def compute_evaluation_loss(
        self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType
    ) -> EvaluationLosses: ...


@runtime_checkable
class AdaptiveDriftConstrainedProtocol(FlexibleClientProtocol, Protocol):
    """What ``AdaptiveDriftConstrainedMixin`` adds to / expects from the client (adaptive_drift_constrained.py:23-32)."""

    loss_for_adaptation: float
    drift_penalty_tensors: list[torch.Tensor] | None
    drift_penalty_weight: float | None
    penalty_loss_function: Any
    parameter_exchanger: Any

    def compute_penalty_loss(self) -> torch.Tensor: ...

    def setup_client_and_return_all_model_parameters(self, config: Config) -> NDArrays: ...


@runtime_checkable
class DittoPersonalizedProtocol(AdaptiveDriftConstrainedProtocol, Protocol):
    """(personalized/ditto.py:30-44)"""

    global_model: nn.Module | None
    optimizer_keys: list[str]

    def get_global_model(self, config: Config) -> nn.Module: ...

    def _copy_optimizer_with_new_params(self, original_optimizer: Optimizer) -> Optimizer: ...

    def set_initial_global_tensors(self) -> None: ...

    def safe_global_model(self) -> nn.Module: ...


@runtime_checkable
class MrMtlPersonalizedProtocol(AdaptiveDriftConstrainedProtocol, Protocol):
    """(personalized/mr_mtl.py:27-32)"""

    initial_global_model: nn.Module | None
    initial_global_tensors: list[torch.Tensor]

    def get_global_model(self, config: Config) -> nn.Module: ...
