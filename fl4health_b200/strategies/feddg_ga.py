"""FedDG-GA: federated domain generalisation with generalisation adjustment (Zhang et al. 2023).

Parity: ``fl4health/strategies/feddg_ga.py:98-477``.  Per-client aggregation weights (keyed by the *stable* client id)
start uniform and move with the gap between the global model's evaluation metric and the local model's post-fit
validation metric: ``a_i += signal * step_r * (gap_i - mean gap) / max|gap - mean gap|``, clipped to [0, 1] and
re-normalised; the step size decays linearly over rounds.  Requires ``evaluate_after_fit`` and
``pack_losses_with_val_metrics`` in the configs and a ``FixedSamplingClientManager`` so fit and evaluate hit the same
clients.  Aggregation goes through ``weighted_combine`` (fused flat kernel / one collective when arena-backed).
"""

from __future__ import annotations

from collections.abc import Callable
from enum import Enum
from logging import INFO, WARNING
from typing import Any

import numpy as np

from fl4health_b200.client_managers.fixed_sampling_client_manager import FixedSamplingClientManager
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
)
from fl4health_b200.servers.client_manager import ClientManager
from fl4health_b200.servers.client_proxy import ClientProxy
from fl4health_b200.strategies.aggregate_utils import weighted_combine
from fl4health_b200.strategies.fedavg import FedAvg
from fl4health_b200.utils.functions import decode_and_pseudo_sort_results


class SignalForTypeError(Exception):
    """Raised when a default signal is requested for a CUSTOM fairness metric."""


class FairnessMetricType(Enum):
    ACCURACY = "val - prediction - accuracy"
    LOSS = "val - checkpoint"
    CUSTOM = "custom"

    @classmethod
    def signal_for_type(cls, fairness_metric_type: FairnessMetricType) -> float:
        if fairness_metric_type == FairnessMetricType.ACCURACY:
            return -1.0
        if fairness_metric_type == FairnessMetricType.LOSS:
            return 1.0
        raise SignalForTypeError("This function should not be called with CUSTOM type.")


class FairnessMetric:
    def __init__(self, metric_type: FairnessMetricType, metric_name: str | None = None, signal: float | None = None) -> None:
        self.metric_type = metric_type
        if metric_type is FairnessMetricType.CUSTOM:
            assert metric_name is not None and signal is not None
            self.metric_name, self.signal = metric_name, signal
        else:
            self.metric_name = metric_name if metric_name is not None else metric_type.value
            self.signal = signal if signal is not None else FairnessMetricType.signal_for_type(metric_type)

    def __str__(self) -> str:
        return f"Metric Type: {self.metric_type}, Metric Name: '{self.metric_name}', Signal: {self.signal}"


class FedDgGa(FedAvg):
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
        initial_parameters: Parameters | None = None,
        fit_metrics_aggregation_fn: MetricsAggregationFn | None = None,
        evaluate_metrics_aggregation_fn: MetricsAggregationFn | None = None,
        fairness_metric: FairnessMetric | None = None,
        adjustment_weight_step_size: float = 0.2,
    ) -> None:
        super().__init__(
            fraction_fit=1.0, fraction_evaluate=1.0, min_fit_clients=min_fit_clients,
            min_evaluate_clients=min_evaluate_clients, min_available_clients=min_available_clients,
            evaluate_fn=evaluate_fn, on_fit_config_fn=on_fit_config_fn, on_evaluate_config_fn=on_evaluate_config_fn,
            accept_failures=accept_failures, initial_parameters=initial_parameters,
            fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
            evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn,
        )
        self.fairness_metric = fairness_metric if fairness_metric is not None else FairnessMetric(FairnessMetricType.LOSS)
        self.adjustment_weight_step_size = adjustment_weight_step_size
        assert 0 < adjustment_weight_step_size < 1, f"adjustment_weight_step_size has to be between 0 and 1 ({adjustment_weight_step_size})"
        log(INFO, f"FedDG-GA Strategy initialized with weight_step_size of {adjustment_weight_step_size} and {self.fairness_metric}")
        self.train_metrics: dict[str, dict[str, Scalar]] = {}
        self.evaluation_metrics: dict[str, dict[str, Scalar]] = {}
        self.num_rounds: int | None = None
        self.initial_adjustment_weight: float | None = None
        self.adjustment_weights: dict[str, float] = {}

    def configure_fit(self, server_round: int, parameters: Parameters, client_manager: ClientManager) -> list[tuple[ClientProxy, FitIns]]:
        assert isinstance(client_manager, FixedSamplingClientManager), f"Client manager is not of type FixedSamplingClientManager: {type(client_manager)}"
        client_manager.reset_sample()
        client_fit_ins = super().configure_fit(server_round, parameters, client_manager)
        self.initial_adjustment_weight = 1.0 / len(client_fit_ins)
        assert self.on_fit_config_fn is not None, "on_fit_config_fn must be specified"
        config = self.on_fit_config_fn(server_round)
        assert config.get("evaluate_after_fit") is True, "evaluate_after_fit must be present and set to True"
        assert config.get("pack_losses_with_val_metrics") is True, "pack_losses_with_val_metrics must be present and True"
        assert isinstance(config.get("n_server_rounds"), int), "n_server_rounds must be specified as an integer"
        n_server_rounds = config["n_server_rounds"]
        if self.num_rounds is None:
            self.num_rounds = n_server_rounds  # type: ignore[assignment]
        else:
            assert n_server_rounds == self.num_rounds, f"n_server_rounds changed from {self.num_rounds} to {n_server_rounds}"
        return client_fit_ins

    def configure_evaluate(self, server_round: int, parameters: Parameters, client_manager: ClientManager) -> list[tuple[ClientProxy, EvaluateIns]]:
        assert isinstance(client_manager, FixedSamplingClientManager)
        client_evaluate_ins = super().configure_evaluate(server_round, parameters, client_manager)
        assert self.on_evaluate_config_fn is not None, "on_evaluate_config_fn must be specified"
        assert self.on_evaluate_config_fn(server_round).get("pack_losses_with_val_metrics") is True
        return client_evaluate_ins

    def aggregate_fit(self, server_round: int, results: list[tuple[ClientProxy, FitRes]], failures: list[Any]) -> tuple[Parameters | None, dict[str, Scalar]]:
        if not results or (not self.accept_failures and failures):
            return None, {}
        metrics = self._aggregate_fit_metrics(server_round, results)
        self.train_metrics = {proxy.cid: res.metrics for proxy, res in results}
        return ndarrays_to_parameters(self.weight_and_aggregate_results(results)), metrics

    def aggregate_evaluate(self, server_round: int, results: list[tuple[ClientProxy, EvaluateRes]], failures: list[Any]) -> tuple[float | None, dict[str, Scalar]]:
        loss_aggregated, metrics_aggregated = super().aggregate_evaluate(server_round, results, failures)
        self.evaluation_metrics = {}
        for proxy, res in results:
            assert FairnessMetricType.LOSS.value in res.metrics
            self.evaluation_metrics[proxy.cid] = res.metrics
        log(INFO, "Updating the Generalization Adjustment Weights")
        self.update_weights_by_ga(server_round, [proxy.cid for proxy, _ in results])
        return loss_aggregated, metrics_aggregated

    def weight_and_aggregate_results(self, results: list[tuple[ClientProxy, FitRes]]) -> NDArrays:
        decoded = decode_and_pseudo_sort_results(results, materialize=False)
        arrays, coefficients = [], []
        for proxy, weights, _ in decoded:
            if proxy.cid not in self.adjustment_weights:
                assert self.initial_adjustment_weight is not None
                self.adjustment_weights[proxy.cid] = self.initial_adjustment_weight
            arrays.append(weights)
            coefficients.append(self.adjustment_weights[proxy.cid])
        log(INFO, f"Current adjustment weights by Client ID (CID) are {self.adjustment_weights}")
        return weighted_combine(arrays, coefficients)

    def update_weights_by_ga(self, server_round: int, cids: list[str]) -> None:
        name = self.fairness_metric.metric_name
        gaps = []
        for cid in cids:
            assert cid in self.train_metrics and cid in self.evaluation_metrics, f"{cid} missing from fit or evaluate metrics"
            global_value, local_value = self.evaluation_metrics[cid][name], self.train_metrics[cid][name]
            assert isinstance(global_value, float) and isinstance(local_value, float)
            gaps.append(global_value - local_value)
        gaps_arr = np.array(gaps)
        centered = gaps_arr - gaps_arr.mean()
        max_dev = np.max(np.abs(centered))
        if max_dev == 0:
            log(WARNING, f"Max variance in generalization gap is 0. Adjustment weights will remain the same. Gaps: {gaps}")
            normalized = np.zeros_like(gaps_arr)
        else:
            normalized = centered * self.get_current_weight_step_size(server_round) / max_dev
        total = 0.0
        for cid, delta in zip(cids, normalized):
            self.adjustment_weights[cid] = float(np.clip(self.adjustment_weights[cid] + self.fairness_metric.signal * delta, 0.0, 1.0))
            total += self.adjustment_weights[cid]
        for cid in cids:
            self.adjustment_weights[cid] /= total
        log(INFO, f"New Generalization Adjustment Weights by Client ID (CID) are {self.adjustment_weights}")

    def get_current_weight_step_size(self, server_round: int) -> float:
        assert self.num_rounds is not None
        decay = self.adjustment_weight_step_size / self.num_rounds
        return self.adjustment_weight_step_size - (server_round - 1) * decay
