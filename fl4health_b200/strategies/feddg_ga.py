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


class GeneralizationAdjustment:
    """The per-client aggregation weights of FedDG-GA and their update rule, separate from the strategy plumbing.

    ``weights`` maps the stable client id to ``a_i``.  ``step(round)`` is the linearly decaying step size
    ``d (1 - (r - 1) / R)``; ``update`` moves every weight by ``signal * step * (gap_i - mean gap) / max |gap - mean gap|``,
    clips to [0, 1] and renormalises over the clients that took part."""

    def __init__(self, step_size: float, signal: float) -> None:
        self.step_size, self.signal = step_size, signal
        self.weights: dict[str, float] = {}
        self.total_rounds: int | None = None

    def step(self, server_round: int) -> float:
        assert self.total_rounds is not None
        return self.step_size * (1.0 - (server_round - 1) / self.total_rounds)

    def weight_of(self, cid: str, default: float) -> float:
        return self.weights.setdefault(cid, default)

    def update(self, server_round: int, gaps: dict[str, float]) -> None:
        cids = list(gaps)
        centred = np.array([gaps[cid] for cid in cids], dtype=float)
        centred -= centred.mean()
        spread = float(np.abs(centred).max())
        if spread == 0:
            log(WARNING, "Max variance in generalization gap is 0. Adjustment weights will remain the same. "
                         f"Gaps: {[gaps[cid] for cid in cids]}")
            moves = np.zeros_like(centred)
        else:
            moves = centred * (self.signal * self.step(server_round) / spread)
        moved = np.clip(np.array([self.weights[cid] for cid in cids]) + moves, 0.0, 1.0)
        moved /= moved.sum()
        self.weights.update({cid: float(value) for cid, value in zip(cids, moved)})


_REQUIRED_FIT_FLAGS = {
    "evaluate_after_fit": "evaluate_after_fit must be present and set to True",
    "pack_losses_with_val_metrics": "pack_losses_with_val_metrics must be present and True",
}


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
        assert 0 < adjustment_weight_step_size < 1, f"adjustment_weight_step_size has to be between 0 and 1 ({adjustment_weight_step_size})"
        super().__init__(
            fraction_fit=1.0, fraction_evaluate=1.0, min_fit_clients=min_fit_clients,
            min_evaluate_clients=min_evaluate_clients, min_available_clients=min_available_clients,
            evaluate_fn=evaluate_fn, on_fit_config_fn=on_fit_config_fn, on_evaluate_config_fn=on_evaluate_config_fn,
            accept_failures=accept_failures, initial_parameters=initial_parameters,
            fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
            evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn,
        )
        self.fairness_metric = fairness_metric or FairnessMetric(FairnessMetricType.LOSS)
        self.adjustment_weight_step_size = adjustment_weight_step_size
        self._adjustment = GeneralizationAdjustment(adjustment_weight_step_size, self.fairness_metric.signal)
        log(INFO, f"FedDG-GA Strategy initialized with weight_step_size of {adjustment_weight_step_size} and {self.fairness_metric}")
        self.train_metrics: dict[str, dict[str, Scalar]] = {}
        self.evaluation_metrics: dict[str, dict[str, Scalar]] = {}
        self.initial_adjustment_weight: float | None = None

    # the reference's attribute names, as views on the adjustment state
    @property
    def adjustment_weights(self) -> dict[str, float]:
        return self._adjustment.weights

    @adjustment_weights.setter
    def adjustment_weights(self, weights: dict[str, float]) -> None:
        self._adjustment.weights = weights

    @property
    def num_rounds(self) -> int | None:
        return self._adjustment.total_rounds

    @num_rounds.setter
    def num_rounds(self, value: int | None) -> None:
        self._adjustment.total_rounds = value

    # ---- round configuration: fit and evaluate must hit the SAME cohort, and clients must report what GA needs ----
    def configure_fit(self, server_round: int, parameters: Parameters, client_manager: ClientManager) -> list[tuple[ClientProxy, FitIns]]:
        assert isinstance(client_manager, FixedSamplingClientManager), f"Client manager is not of type FixedSamplingClientManager: {type(client_manager)}"
        assert self.on_fit_config_fn is not None, "on_fit_config_fn must be specified"
        config = self.on_fit_config_fn(server_round)
        for flag, complaint in _REQUIRED_FIT_FLAGS.items():
            assert config.get(flag) is True, complaint
        declared_rounds = config.get("n_server_rounds")
        assert isinstance(declared_rounds, int), "n_server_rounds must be specified as an integer"
        assert self.num_rounds in (None, declared_rounds), f"n_server_rounds changed from {self.num_rounds} to {declared_rounds}"
        self.num_rounds = declared_rounds
        client_manager.reset_sample()  # a fresh cohort for this round; evaluate re-uses it
        instructions = super().configure_fit(server_round, parameters, client_manager)
        self.initial_adjustment_weight = 1.0 / len(instructions)
        return instructions

    def configure_evaluate(self, server_round: int, parameters: Parameters, client_manager: ClientManager) -> list[tuple[ClientProxy, EvaluateIns]]:
        assert isinstance(client_manager, FixedSamplingClientManager)
        assert self.on_evaluate_config_fn is not None, "on_evaluate_config_fn must be specified"
        assert self.on_evaluate_config_fn(server_round).get("pack_losses_with_val_metrics") is True
        return super().configure_evaluate(server_round, parameters, client_manager)

    # ---- aggregation ----------------------------------------------------------------------------------------------
    def aggregate_fit(self, server_round: int, results: list[tuple[ClientProxy, FitRes]], failures: list[Any]) -> tuple[Parameters | None, dict[str, Scalar]]:
        if not results or (not self.accept_failures and failures):
            return None, {}
        self.train_metrics = {proxy.cid: res.metrics for proxy, res in results}
        fit_metrics = self._aggregate_fit_metrics(server_round, results)
        return ndarrays_to_parameters(self.weight_and_aggregate_results(results)), fit_metrics

    def weight_and_aggregate_results(self, results: list[tuple[ClientProxy, FitRes]]) -> NDArrays:
        assert self.initial_adjustment_weight is not None
        ordered = decode_and_pseudo_sort_results(results, materialize=False)
        coefficients = [self._adjustment.weight_of(proxy.cid, self.initial_adjustment_weight) for proxy, _, _ in ordered]
        log(INFO, f"Current adjustment weights by Client ID (CID) are {self.adjustment_weights}")
        return weighted_combine([weights for _, weights, _ in ordered], coefficients)

    def aggregate_evaluate(self, server_round: int, results: list[tuple[ClientProxy, EvaluateRes]], failures: list[Any]) -> tuple[float | None, dict[str, Scalar]]:
        aggregated = super().aggregate_evaluate(server_round, results, failures)
        assert all(FairnessMetricType.LOSS.value in res.metrics for _, res in results)
        self.evaluation_metrics = {proxy.cid: res.metrics for proxy, res in results}
        log(INFO, "Updating the Generalization Adjustment Weights")
        self.update_weights_by_ga(server_round, [proxy.cid for proxy, _ in results])
        return aggregated

    def update_weights_by_ga(self, server_round: int, cids: list[str]) -> None:
        """Generalisation gap per client = metric of the GLOBAL model on the client's validation data (evaluate round)
        minus the metric of its LOCAL model right after fitting (packed into the fit metrics)."""
        name = self.fairness_metric.metric_name
        gaps: dict[str, float] = {}
        for cid in cids:
            assert cid in self.train_metrics and cid in self.evaluation_metrics, f"{cid} missing from fit or evaluate metrics"
            after_aggregation, after_local_fit = self.evaluation_metrics[cid][name], self.train_metrics[cid][name]
            assert isinstance(after_aggregation, float) and isinstance(after_local_fit, float)
            gaps[cid] = after_aggregation - after_local_fit
        self._adjustment.update(server_round, gaps)
        log(INFO, f"New Generalization Adjustment Weights by Client ID (CID) are {self.adjustment_weights}")

    def get_current_weight_step_size(self, server_round: int) -> float:
        return self._adjustment.step(server_round)
