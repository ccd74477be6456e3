"""FedAvg + server-side adaptation of the drift-penalty weight mu (FedProx / Ditto / MR-MTL).

Parity: ``fl4health/strategies/fedavg_with_adaptive_constraint.py:12-232``.  Wire format: ``weights ++ [mu]`` to the
clients, ``weights ++ [train_loss]`` back.  mu rule: loss did not increase for ``loss_weight_patience`` consecutive
rounds -> ``mu -= delta`` (floored at 0); loss increased -> ``mu += delta``.
"""

from __future__ import annotations

from typing import Any

import numpy as np

from fl4health_b200.common.typing import FitRes, Parameters, Scalar, ndarrays_to_parameters
from fl4health_b200.parameter_exchange.parameter_packer import ParameterPackerAdaptiveConstraint
from fl4health_b200.servers.client_proxy import ClientProxy
from fl4health_b200.strategies.adaptive_weight import LossDrivenWeight
from fl4health_b200.strategies.aggregate_utils import aggregate_losses, aggregate_results
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg
from fl4health_b200.utils.functions import decode_and_pseudo_sort_results


class FedAvgWithAdaptiveConstraint(BasicFedAvg):
    def __init__(
        self,
        *,
        initial_parameters: Parameters | None,
        initial_loss_weight: float = 1.0,
        adapt_loss_weight: bool = False,
        loss_weight_delta: float = 0.1,
        loss_weight_patience: int = 5,
        weighted_train_losses: bool = False,
        **fedavg_options: Any,
    ) -> None:
        """``fedavg_options``: every ``BasicFedAvg`` keyword (``fraction_fit``, ``min_fit_clients``, ``evaluate_fn``,
        ``on_fit_config_fn``, ``weighted_aggregation``, ``weighted_eval_losses``, metric aggregation functions ...)."""
        self._weight = LossDrivenWeight(initial_loss_weight, adapt_loss_weight, loss_weight_delta, loss_weight_patience)
        if initial_parameters:
            self.add_auxiliary_information(initial_parameters)
        super().__init__(initial_parameters=initial_parameters, **fedavg_options)
        self.parameter_packer = ParameterPackerAdaptiveConstraint()
        self.weighted_train_losses = weighted_train_losses

    # the reference's attribute names, served by the controller
    loss_weight = property(lambda self: self._weight.value, lambda self, value: setattr(self._weight, "value", value))
    adapt_loss_weight = property(lambda self: self._weight.adaptive)
    loss_weight_delta = property(lambda self: self._weight.delta)
    loss_weight_patience = property(lambda self: self._weight.patience)
    loss_weight_patience_counter = property(lambda self: self._weight.calm_rounds)
    previous_loss = property(lambda self: self._weight.last_loss)

    def add_auxiliary_information(self, original_parameters: Parameters) -> None:
        """Append the current mu to a parameter set (used on client-initialised parameters too)."""
        original_parameters.tensors.append(np.array(self._weight.value))

    def _maybe_update_constraint_weight_param(self, loss: float) -> None:
        self._weight.observe(loss)

    def aggregate_fit(
        self, server_round: int, results: list[tuple[ClientProxy, FitRes]], failures: list[Any]
    ) -> tuple[Parameters | None, dict[str, Scalar]]:
        if not results or (failures and not self.accept_failures):
            return None, {}
        unpacked = [(self.parameter_packer.unpack_parameters(packed), count)
                    for _, packed, count in decode_and_pseudo_sort_results(results, materialize=False)]
        merged = aggregate_results([(weights, count) for (weights, _), count in unpacked], self.weighted_aggregation)
        round_loss = aggregate_losses([(count, loss) for (_, loss), count in unpacked], self.weighted_train_losses)
        self._maybe_update_constraint_weight_param(round_loss)
        outgoing = self.parameter_packer.pack_parameters(merged, self._weight.value)
        return ndarrays_to_parameters(outgoing), self._aggregate_fit_metrics(server_round, results)
