"""FedAvg with flexible sampling and optional uniform weighting.

Parity: ``fl4health/strategies/basic_fedavg.py:29-400``: ``weighted_aggregation`` / ``weighted_eval_losses`` flags,
sampling through fraction-based client managers, ``configure_poll``, ``add_auxiliary_information`` hook.
"""

from __future__ import annotations

from collections.abc import Callable
from logging import INFO, WARNING
from typing import Any

from torch import nn

from fl4health_b200.client_managers.base_sampling_manager import BaseFractionSamplingManager
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import (
    EvaluateIns,
    EvaluateRes,
    FitIns,
    FitRes,
    GetPropertiesIns,
    MetricsAggregationFn,
    NDArrays,
    Parameters,
    Scalar,
    ndarrays_to_parameters,
)
from fl4health_b200.servers.client_manager import ClientManager
from fl4health_b200.servers.client_proxy import ClientProxy
from fl4health_b200.strategies.aggregate_utils import aggregate_losses, aggregate_results
from fl4health_b200.strategies.fedavg import FedAvg
from fl4health_b200.strategies.strategy_with_poll import StrategyWithPolling
from fl4health_b200.utils.functions import decode_and_pseudo_sort_results


class BasicFedAvg(FedAvg, StrategyWithPolling):
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
        weighted_aggregation: bool = True,
        weighted_eval_losses: bool = True,
    ) -> None:
        super().__init__(
            fraction_fit=fraction_fit, fraction_evaluate=fraction_evaluate, min_fit_clients=min_fit_clients,
            min_evaluate_clients=min_evaluate_clients, min_available_clients=min_available_clients,
            evaluate_fn=evaluate_fn, on_fit_config_fn=on_fit_config_fn, on_evaluate_config_fn=on_evaluate_config_fn,
            accept_failures=accept_failures, initial_parameters=initial_parameters,
            fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
            evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn,
        )
        self.weighted_aggregation = weighted_aggregation
        self.weighted_eval_losses = weighted_eval_losses

    def add_auxiliary_information(self, original_parameters: Parameters) -> None:
        """Hook for strategies that append side information to client-initialised parameters (identity here)."""

    def _sample(self, client_manager: ClientManager, fraction: float, counts: tuple[int, int]) -> list[ClientProxy]:
        if isinstance(client_manager, BaseFractionSamplingManager):
            return client_manager.sample_fraction(fraction, self.min_available_clients)
        sample_size, min_num_clients = counts
        return client_manager.sample(num_clients=sample_size, min_num_clients=min_num_clients)

    def configure_fit(
        self, server_round: int, parameters: Parameters, client_manager: ClientManager
    ) -> list[tuple[ClientProxy, FitIns]]:
        config = self.on_fit_config_fn(server_round) if self.on_fit_config_fn is not None else {}
        fit_ins = FitIns(parameters, config)
        clients = self._sample(client_manager, self.fraction_fit, self.num_fit_clients(client_manager.num_available()))
        return [(client, fit_ins) for client in clients]

    def configure_evaluate(
        self, server_round: int, parameters: Parameters, client_manager: ClientManager
    ) -> list[tuple[ClientProxy, EvaluateIns]]:
        if self.fraction_evaluate == 0.0:
            return []
        config = self.on_evaluate_config_fn(server_round) if self.on_evaluate_config_fn is not None else {}
        evaluate_ins = EvaluateIns(parameters, config)
        clients = self._sample(
            client_manager, self.fraction_evaluate, self.num_evaluation_clients(client_manager.num_available())
        )
        return [(client, evaluate_ins) for client in clients]

    def configure_poll(
        self, server_round: int, client_manager: ClientManager
    ) -> list[tuple[ClientProxy, GetPropertiesIns]]:
        config = self.on_fit_config_fn(server_round) if self.on_fit_config_fn is not None else {}
        property_ins = GetPropertiesIns(config)
        if isinstance(client_manager, BaseFractionSamplingManager):
            clients = client_manager.sample_all(min_num_clients=self.min_available_clients)
        else:
            clients = client_manager.sample(client_manager.num_available(), min_num_clients=self.min_available_clients)
        return [(client, property_ins) for client in clients]

    def aggregate_fit(
        self, server_round: int, results: list[tuple[ClientProxy, FitRes]], failures: list[Any]
    ) -> tuple[Parameters | None, dict[str, Scalar]]:
        if not results or (not self.accept_failures and failures):
            return None, {}
        decoded = decode_and_pseudo_sort_results(results, materialize=False)
        aggregated = aggregate_results([(arrays, n) for _, arrays, n in decoded], self.weighted_aggregation)
        return ndarrays_to_parameters(aggregated), self._aggregate_fit_metrics(server_round, results)

    def aggregate_evaluate(
        self, server_round: int, results: list[tuple[ClientProxy, EvaluateRes]], failures: list[Any]
    ) -> tuple[float | None, dict[str, Scalar]]:
        if not results or (not self.accept_failures and failures):
            return None, {}
        loss = aggregate_losses([(res.num_examples, res.loss) for _, res in results], self.weighted_eval_losses)
        metrics: dict[str, Scalar] = {}
        if self.evaluate_metrics_aggregation_fn is not None:
            metrics = self.evaluate_metrics_aggregation_fn([(res.num_examples, res.metrics) for _, res in results])
        elif server_round == 1:
            log(WARNING, "No evaluate_metrics_aggregation_fn provided")
        return loss, metrics


class OpacusBasicFedAvg(BasicFedAvg):
    """BasicFedAvg whose initial parameters come from a DP-wrapped model (keys carry the wrapper's ``_module.``
    prefix), so server and clients agree on state-dict naming (parity: ``basic_fedavg.py:323-400``)."""

    def __init__(self, *, model: nn.Module, **kwargs: Any) -> None:
        from fl4health_b200.privacy.dp_engine import GradSampleModule

        assert isinstance(model, GradSampleModule), "Provided model must be a GradSampleModule"
        assert kwargs.get("initial_parameters") is None, "initial_parameters are derived from `model`"
        kwargs["initial_parameters"] = ndarrays_to_parameters([v.detach() for v in model.state_dict().values()])
        log(INFO, "Initial parameters taken from the provided GradSampleModule")
        super().__init__(**kwargs)
