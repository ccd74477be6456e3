"""FedDG-GA combined with the adaptive drift-penalty weight of FedProx/Ditto
(parity: ``fl4health/strategies/feddg_ga_with_adaptive_constraint.py:15-241``): GA-weighted aggregation of the model
weights, loss-driven adaptation of mu, wire format ``weights ++ [mu]`` / ``weights ++ [train_loss]``."""

from __future__ import annotations

from collections.abc import Callable
from typing import Any

import numpy as np

from fl4health_b200.common.typing import (
    FitRes,
    MetricsAggregationFn,
    NDArrays,
    Parameters,
    Scalar,
    ndarrays_to_parameters,
    parameters_to_ndarrays,
)
from fl4health_b200.parameter_exchange.parameter_packer import ParameterPackerAdaptiveConstraint
from fl4health_b200.servers.client_proxy import ClientProxy
from fl4health_b200.strategies.aggregate_utils import aggregate_losses
from fl4health_b200.strategies.fedavg_with_adaptive_constraint import FedAvgWithAdaptiveConstraint
from fl4health_b200.strategies.feddg_ga import FairnessMetric, FedDgGa


class FedDgGaAdaptiveConstraint(FedDgGa):
    def __init__(
        self,
        *,
        min_fit_clients: int = 2,
        min_evaluate_clients: int = 2,
        min_available_clients: int = 2,
        evaluate_fn: Callable[[int, NDArrays, dict[str, Scalar]], tuple[float, dict[str, Scalar]] | None] | None = None,
        on_fit_config_fn: Callable[[int], dict[str, Scalar]] | None = None,
        on_evaluate_config_fn: Callable[[int], dict[str, Scalar]] | None = None,
        accept_failures: bool = True,
        initial_parameters: Parameters,
        fit_metrics_aggregation_fn: MetricsAggregationFn | None = None,
        evaluate_metrics_aggregation_fn: MetricsAggregationFn | None = None,
        initial_loss_weight: float = 1.0,
        adapt_loss_weight: bool = False,
        loss_weight_delta: float = 0.1,
        loss_weight_patience: int = 5,
        weighted_train_losses: bool = False,
        fairness_metric: FairnessMetric | None = None,
        adjustment_weight_step_size: float = 0.2,
    ) -> None:
        self.loss_weight = initial_loss_weight
        self.adapt_loss_weight = adapt_loss_weight
        if adapt_loss_weight:
            self.loss_weight_delta, self.loss_weight_patience, self.loss_weight_patience_counter = loss_weight_delta, loss_weight_patience, 0
        self.previous_loss = float("inf")
        if initial_parameters:
            self.add_auxiliary_information(initial_parameters)
        super().__init__(
            min_fit_clients=min_fit_clients, min_evaluate_clients=min_evaluate_clients,
            min_available_clients=min_available_clients, evaluate_fn=evaluate_fn, on_fit_config_fn=on_fit_config_fn,
            on_evaluate_config_fn=on_evaluate_config_fn, accept_failures=accept_failures,
            initial_parameters=initial_parameters, fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
            evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn, fairness_metric=fairness_metric,
            adjustment_weight_step_size=adjustment_weight_step_size,
        )
        self.parameter_packer = ParameterPackerAdaptiveConstraint()
        self.weighted_train_losses = weighted_train_losses

    def add_auxiliary_information(self, original_parameters: Parameters) -> None:
        original_parameters.tensors.append(np.array(self.loss_weight))

    def aggregate_fit(self, server_round: int, results: list[tuple[ClientProxy, FitRes]], failures: list[Any]) -> tuple[Parameters | None, dict[str, Scalar]]:
        if not results or (not self.accept_failures and failures):
            return None, {}
        losses_and_counts = self._unpack_weights_and_losses(results)
        self._maybe_update_constraint_weight_param(aggregate_losses(losses_and_counts, self.weighted_train_losses))
        metrics = self._aggregate_fit_metrics(server_round, results)
        self.train_metrics = {proxy.cid: res.metrics for proxy, res in results}
        weights = self.weight_and_aggregate_results(results)
        return ndarrays_to_parameters(self.parameter_packer.pack_parameters(weights, self.loss_weight)), metrics

    def _unpack_weights_and_losses(self, results: list[tuple[ClientProxy, FitRes]]) -> list[tuple[int, float]]:
        """Strip the packed train loss from every result in place; return (count, loss) pairs."""
        losses_and_counts = []
        for _, res in results:
            weights, train_loss = self.parameter_packer.unpack_parameters(parameters_to_ndarrays(res.parameters))
            tagged = ndarrays_to_parameters(weights)
            if getattr(weights, "ctx", None) is not None:  # keep SPMD ownership tags
                from fl4health_b200.parallel.spmd import _TaggedParameters

                tagged = _TaggedParameters(weights)
            res.parameters = tagged
            losses_and_counts.append((res.num_examples, train_loss))
        return losses_and_counts

    _maybe_update_constraint_weight_param = FedAvgWithAdaptiveConstraint._maybe_update_constraint_weight_param
