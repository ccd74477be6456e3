"""Federated averaging (role of ``flwr.server.strategy.FedAvg``).

Same constructor keywords and sampling rules (``num_fit_clients = max(int(n * fraction_fit), min_fit_clients)``);
aggregation runs through ``aggregate_utils`` (fused flat kernel when clients are arena-backed).
"""

from __future__ import annotations

from collections.abc import Callable
from logging import WARNING
from typing import Any

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
from fl4health_b200.strategies.aggregate_utils import aggregate_results, weighted_loss_avg
from fl4health_b200.strategies.strategy import Strategy
from fl4health_b200.utils.functions import decode_and_pseudo_sort_results

EvaluateFn = Callable[[int, NDArrays, dict[str, Scalar]], "tuple[float, dict[str, Scalar]] | None"]


class FedAvg(Strategy):
    def __init__(
        self,
        *,
        fraction_fit: float = 1.0,
        fraction_evaluate: float = 1.0,
        min_fit_clients: int = 2,
        min_evaluate_clients: int = 2,
        min_available_clients: int = 2,
        evaluate_fn: EvaluateFn | None = None,
        on_fit_config_fn: Callable[[int], dict[str, Scalar]] | None = None,
        on_evaluate_config_fn: Callable[[int], dict[str, Scalar]] | None = None,
        accept_failures: bool = True,
        initial_parameters: Parameters | None = None,
        fit_metrics_aggregation_fn: MetricsAggregationFn | None = None,
        evaluate_metrics_aggregation_fn: MetricsAggregationFn | None = None,
        inplace: bool = True,
    ) -> None:
        if min_fit_clients > min_available_clients or min_evaluate_clients > min_available_clients:
            log(WARNING, "min_fit_clients / min_evaluate_clients exceed min_available_clients.")
        self.fraction_fit = fraction_fit
        self.fraction_evaluate = fraction_evaluate
        self.min_fit_clients = min_fit_clients
        self.min_evaluate_clients = min_evaluate_clients
        self.min_available_clients = min_available_clients
        self.evaluate_fn = evaluate_fn
        self.on_fit_config_fn = on_fit_config_fn
        self.on_evaluate_config_fn = on_evaluate_config_fn
        self.accept_failures = accept_failures
        self.initial_parameters = initial_parameters
        self.fit_metrics_aggregation_fn = fit_metrics_aggregation_fn
        self.evaluate_metrics_aggregation_fn = evaluate_metrics_aggregation_fn
        self.inplace = inplace

    def __repr__(self) -> str:
        return f"{type(self).__name__}(accept_failures={self.accept_failures})"

    def num_fit_clients(self, num_available_clients: int) -> tuple[int, int]:
        return max(int(num_available_clients * self.fraction_fit), self.min_fit_clients), self.min_available_clients

    def num_evaluation_clients(self, num_available_clients: int) -> tuple[int, int]:
        return (
            max(int(num_available_clients * self.fraction_evaluate), self.min_evaluate_clients),
            self.min_available_clients,
        )

    def initialize_parameters(self, client_manager: ClientManager) -> Parameters | None:
        initial, self.initial_parameters = self.initial_parameters, None  # hand over once, then drop the reference
        return initial

    def evaluate(self, server_round: int, parameters: Parameters) -> tuple[float, dict[str, Scalar]] | None:
        if self.evaluate_fn is None:
            return None
        result = self.evaluate_fn(server_round, parameters_to_ndarrays(parameters), {})
        if result is None:
            return None
        loss, metrics = result
        return loss, metrics

    def configure_fit(
        self, server_round: int, parameters: Parameters, client_manager: ClientManager
    ) -> list[tuple[ClientProxy, FitIns]]:
        config = self.on_fit_config_fn(server_round) if self.on_fit_config_fn is not None else {}
        fit_ins = FitIns(parameters, config)
        sample_size, min_num_clients = self.num_fit_clients(client_manager.num_available())
        clients = client_manager.sample(num_clients=sample_size, min_num_clients=min_num_clients)
        return [(client, fit_ins) for client in clients]

    def configure_evaluate(
        self, server_round: int, parameters: Parameters, client_manager: ClientManager
    ) -> list[tuple[ClientProxy, EvaluateIns]]:
        if self.fraction_evaluate == 0.0:
            return []
        config = self.on_evaluate_config_fn(server_round) if self.on_evaluate_config_fn is not None else {}
        evaluate_ins = EvaluateIns(parameters, config)
        sample_size, min_num_clients = self.num_evaluation_clients(client_manager.num_available())
        clients = client_manager.sample(num_clients=sample_size, min_num_clients=min_num_clients)
        return [(client, evaluate_ins) for client in clients]

    def _aggregate_fit_metrics(self, server_round: int, results: list[tuple[ClientProxy, FitRes]]) -> dict[str, Scalar]:
        if self.fit_metrics_aggregation_fn is not None:
            return self.fit_metrics_aggregation_fn([(res.num_examples, res.metrics) for _, res in results])
        if server_round == 1:
            log(WARNING, "No fit_metrics_aggregation_fn provided")
        return {}

    def aggregate_fit(
        self, server_round: int, results: list[tuple[ClientProxy, FitRes]], failures: list[Any]
    ) -> tuple[Parameters | None, dict[str, Scalar]]:
        if not results:
            return None, {}
        if not self.accept_failures and failures:
            return None, {}
        decoded = decode_and_pseudo_sort_results(results, materialize=False)
        aggregated = aggregate_results([(arrays, n) for _, arrays, n in decoded], weighted=True)
        return ndarrays_to_parameters(aggregated), self._aggregate_fit_metrics(server_round, results)

    def aggregate_evaluate(
        self, server_round: int, results: list[tuple[ClientProxy, EvaluateRes]], failures: list[Any]
    ) -> tuple[float | None, dict[str, Scalar]]:
        if not results:
            return None, {}
        if not self.accept_failures and failures:
            return None, {}
        loss = weighted_loss_avg([(res.num_examples, res.loss) for _, res in results])
        metrics: dict[str, Scalar] = {}
        if self.evaluate_metrics_aggregation_fn is not None:
            metrics = self.evaluate_metrics_aggregation_fn([(res.num_examples, res.metrics) for _, res in results])
        elif server_round == 1:
            log(WARNING, "No evaluate_metrics_aggregation_fn provided")
        return loss, metrics
