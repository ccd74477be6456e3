"""FedAvg + server-side adaptation of the drift-penalty weight mu (FedProx / Ditto / MR-MTL).

Parity: ``fl4health/strategies/fedavg_with_adaptive_constraint.py:12-232``.  Wire format: ``weights ++ [mu]`` to the
clients, ``weights ++ [train_loss]`` back.  mu rule: loss did not increase for ``loss_weight_patience`` consecutive
rounds -> ``mu -= delta`` (floored at 0); loss increased -> ``mu += delta``.
"""

from __future__ import annotations

from collections.abc import Callable
from logging import INFO
from typing import Any

import numpy as np

from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import FitRes, MetricsAggregationFn, NDArrays, Parameters, Scalar, ndarrays_to_parameters
from fl4health_b200.parameter_exchange.parameter_packer import ParameterPackerAdaptiveConstraint
from fl4health_b200.servers.client_proxy import ClientProxy
from fl4health_b200.strategies.aggregate_utils import aggregate_losses, aggregate_results
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg
from fl4health_b200.utils.functions import decode_and_pseudo_sort_results


class FedAvgWithAdaptiveConstraint(BasicFedAvg):
    def __init__(
        self,
        *,
        fraction_fit: float = 1.0,
        fraction_evaluate: float = 1.0,
        min_fit_clients: int = 2,
        min_evaluate_clients: int = 2,
        min_available_clients: int = 2,
        evaluate_fn: Callable[[int, NDArrays, dict[str, Scalar]], tuple[float, dict[str, Scalar]] | None] | None = None,
        on_fit_config_fn: Callable[[int], dict[str, Scalar]] | None = None,
        on_evaluate_config_fn: Callable[[int], dict[str, Scalar]] | None = None,
        accept_failures: bool = True,
        initial_parameters: Parameters | None,
        fit_metrics_aggregation_fn: MetricsAggregationFn | None = None,
        evaluate_metrics_aggregation_fn: MetricsAggregationFn | None = None,
        initial_loss_weight: float = 1.0,
        adapt_loss_weight: bool = False,
        loss_weight_delta: float = 0.1,
        loss_weight_patience: int = 5,
        weighted_aggregation: bool = True,
        weighted_eval_losses: bool = True,
        weighted_train_losses: bool = False,
    ) -> None:
        self.loss_weight = initial_loss_weight
        self.adapt_loss_weight = adapt_loss_weight
        if adapt_loss_weight:
            self.loss_weight_delta = loss_weight_delta
            self.loss_weight_patience = loss_weight_patience
            self.loss_weight_patience_counter = 0
        self.previous_loss = float("inf")
        if initial_parameters:
            self.add_auxiliary_information(initial_parameters)
        super().__init__(
            fraction_fit=fraction_fit, fraction_evaluate=fraction_evaluate, min_fit_clients=min_fit_clients,
            min_evaluate_clients=min_evaluate_clients, min_available_clients=min_available_clients,
            evaluate_fn=evaluate_fn, on_fit_config_fn=on_fit_config_fn, on_evaluate_config_fn=on_evaluate_config_fn,
            accept_failures=accept_failures, initial_parameters=initial_parameters,
            fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
            evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn,
            weighted_aggregation=weighted_aggregation, weighted_eval_losses=weighted_eval_losses,
        )
        self.parameter_packer = ParameterPackerAdaptiveConstraint()
        self.weighted_train_losses = weighted_train_losses

    def add_auxiliary_information(self, original_parameters: Parameters) -> None:
        """Append the current mu to a parameter set (used on client-initialised parameters too)."""
        original_parameters.tensors.append(np.array(self.loss_weight))

    def aggregate_fit(
        self, server_round: int, results: list[tuple[ClientProxy, FitRes]], failures: list[Any]
    ) -> tuple[Parameters | None, dict[str, Scalar]]:
        if not results or (not self.accept_failures and failures):
            return None, {}
        weights_and_counts: list[tuple[NDArrays, int]] = []
        losses_and_counts: list[tuple[int, float]] = []
        for _, packed, sample_count in decode_and_pseudo_sort_results(results, materialize=False):
            weights, train_loss = self.parameter_packer.unpack_parameters(packed)
            weights_and_counts.append((weights, sample_count))
            losses_and_counts.append((sample_count, train_loss))
        weights_aggregated = aggregate_results(weights_and_counts, self.weighted_aggregation)
        train_loss_aggregated = aggregate_losses(losses_and_counts, self.weighted_train_losses)
        self._maybe_update_constraint_weight_param(train_loss_aggregated)
        packed_out = self.parameter_packer.pack_parameters(weights_aggregated, self.loss_weight)
        return ndarrays_to_parameters(packed_out), self._aggregate_fit_metrics(server_round, results)

    def _maybe_update_constraint_weight_param(self, loss: float) -> None:
        if self.adapt_loss_weight:
            if loss <= self.previous_loss:
                self.loss_weight_patience_counter += 1
                if self.loss_weight_patience_counter == self.loss_weight_patience:
                    self.loss_weight = max(0.0, self.loss_weight - self.loss_weight_delta)
                    self.loss_weight_patience_counter = 0
                    log(INFO, f"Aggregate training loss has dropped {self.loss_weight_patience} rounds in a row")
                    log(INFO, f"Constraint weight is decreased to {self.loss_weight}")
            else:
                self.loss_weight += self.loss_weight_delta
                self.loss_weight_patience_counter = 0
                log(INFO, f"Aggregate training loss increased this round: Current loss {loss}, Previous loss: {self.previous_loss}")
                log(INFO, f"Constraint weight is increased by {self.loss_weight_delta} to {self.loss_weight}")
        self.previous_loss = loss
