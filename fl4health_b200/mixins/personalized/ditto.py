"""Ditto as a mixin (parity: ``fl4health/mixins/personalized/ditto.py:30-445``): adds a second ("global") model that
is trained without constraint and exchanged with the server, while the client's own model becomes the personal model,
constrained by ``lambda/2 ||w - w_global_init||^2``.  The user's ``get_optimizer`` keeps returning ONE optimizer (for the
personal model); the mixin clones it for the global model."""

from __future__ import annotations

import copy
from logging import INFO, WARNING
from typing import Any

import torch
from torch import nn
from torch.optim import Optimizer

from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config, NDArrays, Scalar
from fl4health_b200.mixins.adaptive_drift_constrained import AdaptiveDriftConstrainedMixin
from fl4health_b200.parameter_exchange.full_exchanger import FullParameterExchanger
from fl4health_b200.utils.config import narrow_dict_type
from fl4health_b200.utils.losses import EvaluationLosses, TrainingLosses
from fl4health_b200.utils.typing import TorchFeatureType, TorchInputType, TorchPredType, TorchTargetType
from fl4health_b200.mixins.core_protocols import DittoPersonalizedProtocol  # noqa: F401  (import-path parity)


class DittoPersonalizedMixin(AdaptiveDriftConstrainedMixin):
    penalty_optimizer_key = "local"
    anchor_from_received_model = False

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        self.global_model: nn.Module | None = None
        super().__init__(*args, **kwargs)

    def safe_global_model(self) -> nn.Module:
        if self.global_model is None:
            raise ValueError("Cannot get global model as it has not yet been set.")
        return self.global_model

    @property
    def optimizer_keys(self) -> list[str]:
        return ["local", "global"]

    def _candidate_modules(self) -> list[nn.Module]:
        return [self.model, self.safe_global_model()]  # type: ignore[attr-defined]

    # ------------------------------------------------------------------------------------------ set-up
    def get_global_model(self, config: Config) -> nn.Module:
        """Same architecture as the personal model (a fresh instance from the user's ``get_model``; clients whose
        ``get_model`` hands out one shared instance — nnU-Net's prepared experiment — get a deep copy)."""
        model = self.get_model(config)  # type: ignore[attr-defined]
        return copy.deepcopy(model) if model is getattr(self, "model", None) else model

    def _copy_optimizer_with_new_params(self, original_optimizer: Optimizer) -> Optimizer:
        """An optimizer of the same class and hyper-parameters (first param group) over the global model."""
        group = original_optimizer.state_dict()["param_groups"][0]
        if "initial_lr" in group:
            initial_lr = group["initial_lr"]
        elif "lr" in original_optimizer.defaults:
            initial_lr = original_optimizer.defaults["lr"]
        else:
            initial_lr = 1e-3
            log(WARNING, "Unable to get the original `lr` for the global optimizer, falling back to `1e-3`.")
        accepted = set(original_optimizer.defaults.keys())
        kwargs = {k: v for k, v in group.items() if k in accepted}
        if type(original_optimizer) is torch.optim.AdamW:
            kwargs.pop("decoupled_weight_decay", None)
        copy = type(original_optimizer)(self.safe_global_model().parameters(), **kwargs)
        for param_group in copy.param_groups:
            param_group["initial_lr"] = initial_lr
        return copy

    def get_optimizer(self, config: Config) -> dict[str, Optimizer]:
        if self.global_model is None:  # needed by the optimizer copy; normally created in setup_client
            self.global_model = self._place_model(self.get_global_model(config))  # type: ignore[attr-defined]
        local = super().get_optimizer(config=config)  # type: ignore[misc]
        if isinstance(local, dict):
            if set(local.keys()) != {"local"}:
                raise ValueError("Ditto mixin expects the wrapped client to define a single optimizer.")
            local = local["local"]
        return {"local": local, "global": self._copy_optimizer_with_new_params(local)}

    def set_optimizer(self, config: Config) -> None:
        optimizers = self.get_optimizer(config)
        assert isinstance(optimizers, dict) and set(optimizers.keys()) == {"global", "local"}
        self.optimizers = optimizers

    def setup_client(self, config: Config) -> None:
        # clients that can only build their model inside their own setup_client (``defer_global_model_creation``) get
        # the global twin created on demand by ``get_optimizer`` instead
        if self.global_model is None and not getattr(self, "defer_global_model_creation", False):
            self.global_model = self._place_model(self.get_global_model(config))  # type: ignore[attr-defined]
            log(INFO, f"Global model set: {type(self.global_model).__name__}")
        super().setup_client(config)  # type: ignore[misc]

    # ------------------------------------------------------------------------------------------ exchange
    def get_parameters(self, config: Config) -> NDArrays:
        if not self.initialized:  # type: ignore[attr-defined]
            return self.setup_client_and_return_all_model_parameters(config)
        if self.initial_parameters_requested(config):  # type: ignore[attr-defined]
            return FullParameterExchanger().push_parameters(self.model, config=config)  # type: ignore[attr-defined]
        assert self.parameter_exchanger is not None  # type: ignore[attr-defined]
        weights = self.parameter_exchanger.push_parameters(self.safe_global_model(), config=config)  # type: ignore[attr-defined]
        return self.parameter_exchanger.pack_parameters(weights, self.loss_for_adaptation)  # type: ignore[attr-defined]

    def set_parameters(self, parameters: NDArrays, config: Config, fitting_round: bool) -> None:
        """Server state goes to the GLOBAL model (and initialises the personal one in round 1)."""
        assert self.parameter_exchanger is not None  # type: ignore[attr-defined]
        server_model_state, self.drift_penalty_weight = self.parameter_exchanger.unpack_parameters(parameters)  # type: ignore[attr-defined]
        log(INFO, f"Lambda weight received from the server: {self.drift_penalty_weight}")
        if narrow_dict_type(config, "current_server_round", int) == 1 and fitting_round:
            log(INFO, "Initializing the global and local models weights for the first time")
            self.initialize_all_model_weights(server_model_state, config)
        else:
            self.parameter_exchanger.pull_parameters(server_model_state, self.safe_global_model(), config)  # type: ignore[attr-defined]

    def initialize_all_model_weights(self, parameters: NDArrays, config: Config) -> None:
        exchanger = FullParameterExchanger()
        exchanger.pull_parameters(parameters, self.model, config)  # type: ignore[attr-defined]
        exchanger.pull_parameters(parameters, self.safe_global_model(), config)

    # ------------------------------------------------------------------------------------------ training
    def set_initial_global_tensors(self) -> None:
        self.drift_penalty_tensors = self.snapshot_drift_anchor(source_model=self.safe_global_model())

    def update_before_train(self, current_server_round: int) -> None:
        self.set_initial_global_tensors()
        self.safe_global_model().train()
        super().update_before_train(current_server_round)

    def train_step(self, input: TorchInputType, target: TorchTargetType) -> tuple[TrainingLosses, TorchPredType]:
        global_model = self.safe_global_model()
        opts = self.optimizers  # type: ignore[attr-defined]
        global_losses, global_preds = self._compute_preds_and_losses(global_model, opts["global"], input, target)  # type: ignore[attr-defined]
        local_losses, local_preds = self._compute_preds_and_losses(self.model, opts["local"], input, target)  # type: ignore[attr-defined]
        local_vanilla = local_losses.backward["backward"].clone()
        global_losses = self._apply_backwards_on_losses_and_take_step(global_model, opts["global"], global_losses)  # type: ignore[attr-defined]
        penalty = self.compute_penalty_loss()
        local_losses.backward["backward"] = local_losses.backward["backward"] + penalty
        local_losses = self._apply_backwards_on_losses_and_take_step(self.model, opts["local"], local_losses)  # type: ignore[attr-defined]
        local_losses.additional_losses = {
            "penalty_loss": penalty.clone(), "local_loss": local_vanilla,
            "global_loss": global_losses.backward["backward"].detach(), "loss_for_adaptation": local_vanilla.clone(),
        }
        return local_losses, _combine_predictions(global_preds, local_preds)

    def val_step(self, input: TorchInputType, target: TorchTargetType) -> tuple[EvaluationLosses, TorchPredType]:
        global_losses, global_preds = self._val_step_with_model(self.safe_global_model(), input, target)  # type: ignore[attr-defined]
        local_losses, local_preds = self._val_step_with_model(self.model, input, target)  # type: ignore[attr-defined]
        losses = EvaluationLosses(
            local_losses.checkpoint,
            additional_losses={"global_loss": global_losses.checkpoint, "local_loss": local_losses.checkpoint},
        )
        return losses, _combine_predictions(global_preds, local_preds)

    def validate(self, include_losses_in_metrics: bool = False) -> tuple[float, dict[str, Scalar]]:
        self.safe_global_model().eval()
        return super().validate(include_losses_in_metrics=include_losses_in_metrics)  # type: ignore[misc]

    def compute_evaluation_loss(
        self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType
    ) -> EvaluationLosses:
        assert self.global_model is not None and not self.global_model.training and not self.model.training  # type: ignore[attr-defined]
        return super().compute_evaluation_loss(preds, features, target)  # type: ignore[misc]


def _combine_predictions(global_preds: TorchPredType, local_preds: TorchPredType) -> TorchPredType:
    combined = {f"global-{k}": v for k, v in global_preds.items()}
    combined.update({f"local-{k}": v for k, v in local_preds.items()})
    return combined
