"""Plain federated averaging (McMahan et al. 2017) with Flower's constructor vocabulary."""

from __future__ import annotations

from collections.abc import Callable
from logging import WARNING

from ...common.logger import log
from ...common.parameter import ndarrays_to_parameters, parameters_to_ndarrays
from ...common.typing import EvaluateIns, EvaluateRes, FitIns, FitRes, MetricsAggregationFn, NDArrays, Parameters, Scalar
from ..client_manager import ClientManager
from ..client_proxy import ClientProxy
from .aggregate import aggregate, weighted_loss_avg
from .strategy import Strategy


class FedAvg(Strategy):
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
        initial_parameters: Parameters | None = None,
        fit_metrics_aggregation_fn: MetricsAggregationFn | None = None,
        evaluate_metrics_aggregation_fn: MetricsAggregationFn | None = None,
        inplace: bool = True,
    ) -> None:
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
        return f"FedAvg(accept_failures={self.accept_failures})"

    def num_fit_clients(self, num_available_clients: int) -> tuple[int, int]:
        return max(int(num_available_clients * self.fraction_fit), self.min_fit_clients), self.min_available_clients

    def num_evaluation_clients(self, num_available_clients: int) -> tuple[int, int]:
        return max(int(num_available_clients * self.fraction_evaluate), self.min_evaluate_clients), self.min_available_clients

    def initialize_parameters(self, client_manager: ClientManager) -> Parameters | None:
        initial, self.initial_parameters = self.initial_parameters, None
        return initial

    def evaluate(self, server_round: int, parameters: Parameters) -> tuple[float, dict[str, Scalar]] | None:
        if self.evaluate_fn is None:
            return None
        outcome = self.evaluate_fn(server_round, parameters_to_ndarrays(parameters), {})
        return None if outcome is None else (outcome[0], outcome[1])

    def configure_fit(self, server_round: int, parameters: Parameters, client_manager: ClientManager) -> list[tuple[ClientProxy, FitIns]]:
        config = self.on_fit_config_fn(server_round) if self.on_fit_config_fn is not None else {}
        ins = FitIns(parameters, config)
        size, floor = self.num_fit_clients(client_manager.num_available())
        return [(client, ins) for client in client_manager.sample(num_clients=size, min_num_clients=floor)]

    def configure_evaluate(self, server_round: int, parameters: Parameters, client_manager: ClientManager) -> list[tuple[ClientProxy, EvaluateIns]]:
        if self.fraction_evaluate == 0.0:
            return []
        config = self.on_evaluate_config_fn(server_round) if self.on_evaluate_config_fn is not None else {}
        ins = EvaluateIns(parameters, config)
        size, floor = self.num_evaluation_clients(client_manager.num_available())
        return [(client, ins) for client in client_manager.sample(num_clients=size, min_num_clients=floor)]

    def aggregate_fit(self, server_round: int, results: list[tuple[ClientProxy, FitRes]],
                      failures: list[tuple[ClientProxy, FitRes] | BaseException]) -> tuple[Parameters | None, dict[str, Scalar]]:
        if not results or (failures and not self.accept_failures):
            return None, {}
        merged = aggregate([(parameters_to_ndarrays(res.parameters), res.num_examples) for _, res in results])
        metrics: dict[str, Scalar] = {}
        if self.fit_metrics_aggregation_fn:
            metrics = self.fit_metrics_aggregation_fn([(res.num_examples, res.metrics) for _, res in results])
        elif server_round == 1:
            log(WARNING, "No fit_metrics_aggregation_fn provided")
        return ndarrays_to_parameters(merged), metrics

    def aggregate_evaluate(self, server_round: int, results: list[tuple[ClientProxy, EvaluateRes]],
                           failures: list[tuple[ClientProxy, EvaluateRes] | BaseException]) -> tuple[float | None, dict[str, Scalar]]:
        if not results or (failures and not self.accept_failures):
            return None, {}
        loss = weighted_loss_avg([(res.num_examples, res.loss) for _, res in results])
        metrics: dict[str, Scalar] = {}
        if self.evaluate_metrics_aggregation_fn:
            metrics = self.evaluate_metrics_aggregation_fn([(res.num_examples, res.metrics) for _, res in results])
        elif server_round == 1:
            log(WARNING, "No evaluate_metrics_aggregation_fn provided")
        return loss, metrics
