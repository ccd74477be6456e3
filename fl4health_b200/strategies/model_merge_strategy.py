"""Strategy for one-shot model merging (parity: ``fl4health/strategies/model_merge_strategy.py:26-282``): sends EMPTY
parameters to the clients, averages the models they send back; evaluation aggregates metrics only."""

from __future__ import annotations

from collections.abc import Callable
from logging import WARNING
from typing import Any

from fl4health_b200.client_managers.base_sampling_manager import BaseFractionSamplingManager
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import (
    EvaluateIns,
    EvaluateRes,
    FitIns,
    FitRes,
    MetricsAggregationFn,
    NDArrays,
    Parameters,
    Scalar,
    ndarrays_to_parameters,
    parameters_to_ndarrays,
)
from fl4health_b200.servers.client_manager import ClientManager
from fl4health_b200.servers.client_proxy import ClientProxy
from fl4health_b200.strategies.aggregate_utils import aggregate_results
from fl4health_b200.strategies.strategy import Strategy
from fl4health_b200.utils.functions import decode_and_pseudo_sort_results


class ModelMergeStrategy(Strategy):
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
        fit_metrics_aggregation_fn: MetricsAggregationFn | None = None,
        evaluate_metrics_aggregation_fn: MetricsAggregationFn | None = None,
        weighted_aggregation: bool = True,
    ) -> None:
        self.fraction_fit, self.fraction_evaluate = fraction_fit, fraction_evaluate
        self.min_fit_clients, self.min_evaluate_clients = min_fit_clients, min_evaluate_clients
        self.min_available_clients = min_available_clients
        self.evaluate_fn = evaluate_fn
        self.on_fit_config_fn, self.on_evaluate_config_fn = on_fit_config_fn, on_evaluate_config_fn
        self.accept_failures = accept_failures
        self.fit_metrics_aggregation_fn = fit_metrics_aggregation_fn
        self.evaluate_metrics_aggregation_fn = evaluate_metrics_aggregation_fn
        self.weighted_aggregation = weighted_aggregation

    def _sample(self, client_manager: ClientManager, fraction: float, minimum: int) -> list[ClientProxy]:
        if isinstance(client_manager, BaseFractionSamplingManager):
            return client_manager.sample_fraction(fraction, self.min_available_clients)
        sample_size = max(int(client_manager.num_available() * fraction), minimum)
        return client_manager.sample(num_clients=sample_size, min_num_clients=self.min_available_clients)

    def configure_fit(self, server_round: int, parameters: Parameters, client_manager: ClientManager) -> list[tuple[ClientProxy, FitIns]]:
        config = self.on_fit_config_fn(server_round) if self.on_fit_config_fn is not None else {}
        fit_ins = FitIns(Parameters([], ""), config)  # clients bring their own weights
        return [(client, fit_ins) for client in self._sample(client_manager, self.fraction_fit, self.min_fit_clients)]

    def configure_evaluate(self, server_round: int, parameters: Parameters, client_manager: ClientManager) -> list[tuple[ClientProxy, EvaluateIns]]:
        if self.fraction_evaluate == 0.0:
            return []
        config = self.on_evaluate_config_fn(server_round) if self.on_evaluate_config_fn is not None else {}
        evaluate_ins = EvaluateIns(parameters, config)
        return [(c, evaluate_ins) for c in self._sample(client_manager, self.fraction_evaluate, self.min_evaluate_clients)]

    def aggregate_fit(self, server_round: int, results: list[tuple[ClientProxy, FitRes]], failures: list[Any]) -> tuple[Parameters | None, dict[str, Scalar]]:
        if not results or (not self.accept_failures and failures):
            return None, {}
        decoded = [(arrays, n) for _, arrays, n in decode_and_pseudo_sort_results(results, materialize=False)]
        merged = aggregate_results(decoded, self.weighted_aggregation)
        metrics: dict[str, Scalar] = {}
        if self.fit_metrics_aggregation_fn:
            metrics = self.fit_metrics_aggregation_fn([(res.num_examples, res.metrics) for _, res in results])
        elif server_round == 1:
            log(WARNING, "No fit_metrics_aggregation_fn provided")
        return ndarrays_to_parameters(merged), metrics

    def aggregate_evaluate(self, server_round: int, results: list[tuple[ClientProxy, EvaluateRes]], failures: list[Any]) -> tuple[float | None, dict[str, Scalar]]:
        if not results or (not self.accept_failures and failures):
            return None, {}
        metrics: dict[str, Scalar] = {}
        if self.evaluate_metrics_aggregation_fn:
            metrics = self.evaluate_metrics_aggregation_fn([(res.num_examples, res.metrics) for _, res in results])
        elif server_round == 1:
            log(WARNING, "No evaluate_metrics_aggregation_fn provided")
        return None, metrics

    def evaluate(self, server_round: int, parameters: Parameters) -> tuple[float, dict[str, Scalar]] | None:
        if self.evaluate_fn is None:
            return None
        result = self.evaluate_fn(server_round, parameters_to_ndarrays(parameters), {})
        return None if result is None else (result[0], result[1])

    def initialize_parameters(self, client_manager: ClientManager) -> Parameters | None:
        return None
