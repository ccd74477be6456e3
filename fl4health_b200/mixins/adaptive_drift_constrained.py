"""FedProx-style adaptive drift constraint as a mixin over any ``FlexibleClient`` (parity:
``fl4health/mixins/adaptive_drift_constrained.py:23-224``).

Wire protocol is identical to ``AdaptiveDriftConstraintClient``: the server ships ``(weights, mu)``, the client
returns ``(weights, vanilla training loss)``.  When the model lives in a flat arena with a fused optimizer the penalty
gradient ``mu (w - w_ref)`` is folded into the optimizer kernel (see ``engine/fused_optim.py``) and only the penalty
VALUE is computed here.
"""

from __future__ import annotations

from logging import INFO
from typing import Any

import torch
from torch import nn

from fl4health_b200.clients.flexible.base import FlexibleClient
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config, NDArrays
from fl4health_b200.engine.fused_optim import _FlatOptimizer
from fl4health_b200.losses.weight_drift_loss import WeightDriftLoss
from fl4health_b200.mixins.base import BaseFlexibleMixin
from fl4health_b200.ops import flat as flat_ops
from fl4health_b200.parallel.arena import arena_of
from fl4health_b200.parameter_exchange.full_exchanger import FullParameterExchanger
from fl4health_b200.parameter_exchange.packing_exchanger import FullParameterExchangerWithPacking
from fl4health_b200.parameter_exchange.parameter_exchanger_base import ParameterExchanger
from fl4health_b200.parameter_exchange.parameter_packer import ParameterPackerAdaptiveConstraint
from fl4health_b200.utils.losses import TrainingLosses
from fl4health_b200.utils.typing import TorchInputType, TorchPredType, TorchTargetType
from fl4health_b200.mixins.core_protocols import AdaptiveDriftConstrainedProtocol  # noqa: F401  (import-path parity)


class AdaptiveDriftConstrainedMixin(BaseFlexibleMixin):
    penalty_optimizer_key = "global"
    # plain FedProx: the anchor w_ref is the model as received this round; personalised mixins set their own anchor
    anchor_from_received_model = True

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        self.loss_for_adaptation = 0.1
        self.drift_penalty_tensors: list[torch.Tensor] | None = None
        self.drift_penalty_weight: float | None = None
        super().__init__(*args, **kwargs)
        self.penalty_loss_function = WeightDriftLoss(self.device)  # type: ignore[attr-defined]

    # ---------------------------------------------------------------------------------------- exchange
    def get_parameter_exchanger(self, config: Config) -> ParameterExchanger:
        return FullParameterExchangerWithPacking(ParameterPackerAdaptiveConstraint())

    def get_parameters(self, config: Config) -> NDArrays:
        if not self.initialized:  # type: ignore[attr-defined]
            return self.setup_client_and_return_all_model_parameters(config)
        if self.initial_parameters_requested(config):  # type: ignore[attr-defined]
            return FullParameterExchanger().push_parameters(self.model, config=config)  # type: ignore[attr-defined]
        assert self.model is not None and self.parameter_exchanger is not None  # type: ignore[attr-defined]
        weights = self.parameter_exchanger.push_parameters(self.model, config=config)  # type: ignore[attr-defined]
        return self.parameter_exchanger.pack_parameters(weights, self.loss_for_adaptation)  # type: ignore[attr-defined]

    def setup_client_and_return_all_model_parameters(self, config: Config) -> NDArrays:
        log(INFO, "Setting up client and providing full model parameters to the server for initialization")
        if not config:
            log(INFO, "Config sent by the server is empty: setup may fail if it needs config entries")
        self.setup_client(config)  # type: ignore[attr-defined]
        return FullParameterExchanger().push_parameters(self.model, config=config)  # type: ignore[attr-defined]

    def set_parameters(self, parameters: NDArrays, config: Config, fitting_round: bool) -> None:
        assert self.model is not None and self.parameter_exchanger is not None  # type: ignore[attr-defined]
        server_model_state, self.drift_penalty_weight = self.parameter_exchanger.unpack_parameters(parameters)  # type: ignore[attr-defined]
        log(INFO, f"Penalty weight received from the server: {self.drift_penalty_weight}")
        super().set_parameters(server_model_state, config, fitting_round)  # type: ignore[misc]

    # ---------------------------------------------------------------------------------------- penalty
    def _constrained_model(self) -> nn.Module:
        return self.model  # type: ignore[attr-defined]

    def snapshot_drift_anchor(self, source_model: nn.Module | None = None) -> list[torch.Tensor]:
        """Anchor tensors ``w_ref`` for the penalty; mirrored into the arena's ``drift_anchor`` region when present so
        the fused optimizer can apply ``mu (w - w_ref)`` in the same pass."""
        constrained = self._constrained_model()
        source = source_model if source_model is not None else constrained
        arena, src_arena = arena_of(constrained), arena_of(source)
        if arena is not None and src_arena is not None and arena.same_layout(src_arena):
            anchor = arena.companion("drift_anchor")
            if src_arena is arena and self.anchor_from_received_model:
                arena.anchor_on_pull = True  # from now on the pull kernel writes the anchor itself
            if src_arena is arena and arena.anchor_fresh:
                arena.anchor_fresh = False   # this pull already produced w_t
            else:
                flat_ops.bcast_unpack(src_arena.flat, w=None, anchor=anchor)
            return [arena.view(name, anchor) for name, _ in constrained.named_parameters()]
        return [p.detach().clone() for p in source.parameters()]

    def _fused_penalty_optimizer(self) -> _FlatOptimizer | None:
        if type(self).compute_penalty_loss is not AdaptiveDriftConstrainedMixin.compute_penalty_loss:
            return None
        optimizer = getattr(self, "optimizers", {}).get(self.penalty_optimizer_key)
        arena = arena_of(self._constrained_model())
        if not isinstance(optimizer, _FlatOptimizer) or arena is None or optimizer.arena is not arena:
            return None
        if "drift_anchor" not in arena.regions or self.drift_penalty_tensors is None:
            return None
        return optimizer

    def compute_penalty_loss(self) -> torch.Tensor:
        assert self.drift_penalty_tensors is not None and self.drift_penalty_weight is not None
        optimizer = self._fused_penalty_optimizer()
        if optimizer is not None:
            arena, anchor = optimizer.arena, optimizer.arena.regions["drift_anchor"]
            optimizer.set_drift_anchor(anchor, self.drift_penalty_weight)
            with torch.no_grad():
                n = arena.trainable_padded
                return (flat_ops.sq_diff_sum(arena.flat[:n], anchor[:n]) * (self.drift_penalty_weight / 2.0)).reshape(())
        return self.penalty_loss_function(self._constrained_model(), self.drift_penalty_tensors, self.drift_penalty_weight)

    # ---------------------------------------------------------------------------------------- step
    def update_before_train(self, current_server_round: int) -> None:
        if self.anchor_from_received_model:
            self.drift_penalty_tensors = self.snapshot_drift_anchor()
        super().update_before_train(current_server_round)  # type: ignore[misc]

    def train_step(self, input: TorchInputType, target: TorchTargetType) -> tuple[TrainingLosses, TorchPredType]:
        optimizer = self.optimizers[self.penalty_optimizer_key]  # type: ignore[attr-defined]
        losses, preds = self._compute_preds_and_losses(self.model, optimizer, input, target)  # type: ignore[attr-defined]
        vanilla = losses.backward["backward"].clone()
        penalty = self.compute_penalty_loss()
        losses.backward["backward"] = losses.backward["backward"] + penalty
        losses = self._apply_backwards_on_losses_and_take_step(self.model, optimizer, losses)  # type: ignore[attr-defined]
        losses.additional_losses = {"penalty_loss": penalty.clone(), "local_loss": vanilla, "loss_for_adaptation": vanilla.clone()}
        return losses, preds

    def update_after_train(self, local_steps: int, loss_dict: dict[str, float], config: Config) -> None:
        assert "loss_for_adaptation" in loss_dict
        self.loss_for_adaptation = loss_dict["loss_for_adaptation"]
        super().update_after_train(local_steps, loss_dict, config)  # type: ignore[misc]


def apply_adaptive_drift_to_client(client_base_type: type[FlexibleClient]) -> type[FlexibleClient]:
    """Dynamically create ``AdaptiveDrift<Client>`` = mixin + the given flexible client class."""
    return type(
        f"AdaptiveDrift{client_base_type.__name__}", (AdaptiveDriftConstrainedMixin, client_base_type),
        {"_dynamically_created": True},
    )
