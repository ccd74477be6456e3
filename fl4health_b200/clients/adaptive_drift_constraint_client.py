"""Base client for drift-penalised training (FedProx, Ditto, MR-MTL).

Parity: ``fl4health/clients/adaptive_drift_constraint_client.py:21-203``: the server ships ``weights ++ [mu]``; the
client trains on ``loss + (mu/2)||w - w_ref||^2`` and returns ``weights ++ [vanilla train loss]`` so the server can
adapt ``mu``.

Engine fast path: with an arena-backed model and a flat fused optimizer the penalty *gradient* ``mu (w - w_ref)`` is
added inside the optimizer kernel and the penalty *value* comes from one flat reduction — the autograd graph never
sees the penalty (identical update, 3·L fewer nodes per step).  Any subclass overriding ``compute_penalty_loss`` or
using a non-translatable optimizer transparently gets the autograd path.
"""

from __future__ import annotations

from logging import INFO
from typing import Any

import torch

from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config, NDArrays
from fl4health_b200.engine.fused_optim import _FlatOptimizer
from fl4health_b200.losses.weight_drift_loss import WeightDriftLoss
from fl4health_b200.ops import flat as flat_ops
from fl4health_b200.parallel.arena import arena_of
from fl4health_b200.parameter_exchange.full_exchanger import FullParameterExchanger
from fl4health_b200.parameter_exchange.packing_exchanger import FullParameterExchangerWithPacking
from fl4health_b200.parameter_exchange.parameter_exchanger_base import ParameterExchanger
from fl4health_b200.parameter_exchange.parameter_packer import ParameterPackerAdaptiveConstraint
from fl4health_b200.utils.losses import TrainingLosses
from fl4health_b200.utils.typing import TorchFeatureType, TorchPredType, TorchTargetType


class AdaptiveDriftConstraintClient(BasicClient):
    """Constructor arguments are ``BasicClient``'s.  Subclasses describe their variant declaratively:

    * ``exchanged_model``       attribute holding the network that travels to / from the server (Ditto: its global twin);
    * ``anchor_model``          attribute whose weights are the drift reference ``w_ref`` at the start of local training
                                (None: no automatic snapshot; "model": the weights just received, FedProx);
    * ``penalty_optimizer_key`` the optimizer stepping the *constrained* network ``self.model`` (absorbs ``mu (w - w_ref)``);
    * ``receives_into``         attribute the aggregate is written to outside the very first fit.
    """

    exchanged_model = "model"
    receives_into = "model"
    anchor_model: str | None = None
    penalty_optimizer_key = "global"

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self.parameter_exchanger: FullParameterExchangerWithPacking[float]
        self.drift_penalty_tensors: list[torch.Tensor] | None = None
        self.drift_penalty_weight: float | None = None
        self.loss_for_adaptation: float = 0.0
        self.penalty_loss_function = WeightDriftLoss(self.device)

    # ---------------------------------------------------------------------------------------- wire format
    def get_parameter_exchanger(self, config: Config) -> ParameterExchanger:
        return FullParameterExchangerWithPacking(ParameterPackerAdaptiveConstraint())

    def get_parameters(self, config: Config) -> NDArrays:
        """``weights(exchanged model) ++ [vanilla training loss]`` — the loss lets the server adapt the penalty weight."""
        if not self.initialized:
            return self.setup_client_and_return_all_model_parameters(config)
        if self.initial_parameters_requested(config):  # already set up by a properties poll: plain model state, unpacked
            return FullParameterExchanger().push_parameters(self.model, config=config)
        outgoing = self.parameter_exchanger.push_parameters(getattr(self, self.exchanged_model), config=config)
        return self.parameter_exchanger.pack_parameters(outgoing, self.loss_for_adaptation)

    def set_parameters(self, parameters: NDArrays, config: Config, fitting_round: bool) -> None:
        """``weights ++ [mu]`` from the server: remember ``mu``, route the weights to where this variant keeps them."""
        aggregate, self.drift_penalty_weight = self.parameter_exchanger.unpack_parameters(parameters)
        log(INFO, f"Penalty weight received from the server: {self.drift_penalty_weight}")
        self._install_aggregate(aggregate, config, fitting_round)

    def _install_aggregate(self, aggregate: NDArrays, config: Config, fitting_round: bool) -> None:
        if self.receives_into == "model":
            BasicClient.set_parameters(self, aggregate, config, fitting_round)
        elif fitting_round and config.get("current_server_round") == 1 and self.exchanged_model != "model" \
                and self.receives_into == self.exchanged_model:
            log(INFO, "Initializing the global and local models weights for the first time")
            self.initialize_all_model_weights(aggregate, config)
        else:
            self.parameter_exchanger.pull_parameters(aggregate, getattr(self, self.receives_into), config)

    def update_before_train(self, current_server_round: int) -> None:
        if self.anchor_model is not None:
            self.drift_penalty_tensors = self.snapshot_drift_anchor(
                source_model=getattr(self, self.anchor_model), constrained_model=self.model)
        super().update_before_train(current_server_round)

    def update_after_train(self, local_steps: int, loss_dict: dict[str, float], config: Config) -> None:
        assert "loss_for_adaptation" in loss_dict
        self.loss_for_adaptation = loss_dict["loss_for_adaptation"]
        super().update_after_train(local_steps, loss_dict, config)

    # ---------------------------------------------------------------------------------------- drift anchor
    def snapshot_drift_anchor(self, source_model: torch.nn.Module | None = None, constrained_model: torch.nn.Module | None = None) -> list[torch.Tensor]:
        """Reference tensors for the penalty = current parameters of ``source_model`` (default: the trained model).

        Arena-backed models snapshot with ONE flat copy into an arena-shaped companion region of the *constrained*
        model; the returned list contains per-parameter views of that region (API-compatible with the reference's list
        of cloned tensors)."""
        constrained = constrained_model if constrained_model is not None else self.model
        source = source_model if source_model is not None else constrained
        dst_arena, src_arena = arena_of(constrained), arena_of(source)
        if dst_arena is not None and src_arena is not None and dst_arena.same_layout(src_arena):
            anchor = dst_arena.companion("drift_anchor")
            if src_arena is dst_arena and getattr(self, "anchor_from_received_model", True):
                dst_arena.anchor_on_pull = True  # from now on the pull kernel writes the anchor itself
            if src_arena is dst_arena and dst_arena.anchor_fresh:
                dst_arena.anchor_fresh = False   # this pull already produced w_t
            else:
                flat_ops.bcast_unpack(src_arena.flat, w=None, anchor=anchor)
            return [dst_arena.view(name, anchor) for name, _ in constrained.named_parameters()]
        return [p.detach().clone() for p in source.parameters()]

    def _fused_penalty_optimizer(self) -> _FlatOptimizer | None:
        """The flat optimizer that can absorb the penalty gradient, if the fast path applies."""
        if type(self).compute_penalty_loss is not AdaptiveDriftConstraintClient.compute_penalty_loss:
            return None
        optimizer = self.optimizers.get(self.penalty_optimizer_key) if hasattr(self, "optimizers") else None
        arena = arena_of(self.model)
        if not isinstance(optimizer, _FlatOptimizer) or arena is None or optimizer.arena is not arena:
            return None
        if "drift_anchor" not in arena.regions or self.drift_penalty_tensors is None:
            return None
        return optimizer

    # ---------------------------------------------------------------------------------------- losses
    def compute_training_loss(
        self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType
    ) -> TrainingLosses:
        task_loss, extras = self.compute_loss_and_additional_losses(preds, features, target)
        penalty = self.compute_penalty_loss()
        recorded = {**(extras or {}), "loss": task_loss.clone(), "loss_for_adaptation": task_loss.clone(),
                    "penalty_loss": penalty.clone()}
        return TrainingLosses(backward=task_loss + penalty, additional_losses=recorded)

    def compute_penalty_loss(self) -> torch.Tensor:
        assert self.drift_penalty_tensors is not None and self.drift_penalty_weight is not None
        optimizer = self._fused_penalty_optimizer()
        if optimizer is not None:
            arena = optimizer.arena
            anchor = arena.regions["drift_anchor"]
            optimizer.set_drift_anchor(anchor, self.drift_penalty_weight)
            with torch.no_grad():  # value only: the gradient mu (w - w_ref) is added inside the optimizer kernel
                n = arena.trainable_padded
                return (flat_ops.sq_diff_sum(arena.flat[:n], anchor[:n]) * (self.drift_penalty_weight / 2.0)).reshape(())
        return self.penalty_loss_function(self.model, self.drift_penalty_tensors, self.drift_penalty_weight)
