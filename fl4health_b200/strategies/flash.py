"""FLASH server optimizer (Panchal et al. 2023; parity: ``fl4health/strategies/flash.py:21-170``): Adam-style server
step with a drift-aware second moment:  ``d <- b3 d + (1-b3)(delta^2 - v)``, ``b3 = |v_prev| / (|delta^2 - v| + |v_prev|)``,
``w <- w + eta * m / (sqrt(v) - d + tau)``."""

from __future__ import annotations

from collections.abc import Callable
from typing import Any

import torch

from fl4health_b200.common.typing import (
    FitRes,
    MetricsAggregationFn,
    NDArrays,
    Parameters,
    Scalar,
    ndarrays_to_parameters,
    parameters_to_ndarrays,
    to_tensor,
)
from fl4health_b200.servers.client_proxy import ClientProxy
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg

EVALUATE_FN_TYPE = Callable[[int, NDArrays, dict[str, Scalar]], tuple[float, dict[str, Scalar]] | None] | None


class Flash(BasicFedAvg):
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
        eta: float = 1e-1,
        eta_l: float = 1e-1,
        beta_1: float = 0.9,
        beta_2: float = 0.99,
        tau: float = 1e-9,
        weighted_aggregation: bool = False,
        weighted_eval_losses: bool = False,
    ) -> None:
        super().__init__(
            fraction_fit=fraction_fit, fraction_evaluate=fraction_evaluate, min_fit_clients=min_fit_clients,
            min_evaluate_clients=min_evaluate_clients, min_available_clients=min_available_clients,
            evaluate_fn=evaluate_fn, on_fit_config_fn=on_fit_config_fn, on_evaluate_config_fn=on_evaluate_config_fn,
            accept_failures=accept_failures, initial_parameters=initial_parameters,
            fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
            evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn,
            weighted_aggregation=weighted_aggregation, weighted_eval_losses=weighted_eval_losses,
        )
        self.eta, self.eta_l, self.tau, self.beta_1, self.beta_2 = eta, eta_l, tau, beta_1, beta_2
        self.current_weights: list[torch.Tensor] = []
        self.m_t: list[torch.Tensor] = []
        self.v_t: list[torch.Tensor] = []
        self.d_t: list[torch.Tensor] = []
        if initial_parameters:
            self._initialise(parameters_to_ndarrays(initial_parameters))

    def __repr__(self) -> str:
        return f"Flash(accept_failures={self.accept_failures})"

    def _initialise(self, weights: NDArrays) -> None:
        self.current_weights = [to_tensor(w).clone() for w in weights]
        self.m_t = [torch.zeros_like(w, dtype=torch.float32) for w in self.current_weights]
        self.v_t = [torch.zeros_like(w, dtype=torch.float32) for w in self.current_weights]
        self.d_t = [torch.zeros_like(w, dtype=torch.float32) for w in self.current_weights]

    def add_auxiliary_information(self, original_parameters: Parameters) -> None:
        """Client-initialised parameters become the server's starting point (nothing is packed in)."""
        self._initialise(parameters_to_ndarrays(original_parameters))

    def _update_parameters(self, delta_t: list[torch.Tensor]) -> None:
        for i, delta in enumerate(delta_t):
            m_prev, v_prev, d_prev = self.m_t[i], self.v_t[i], self.d_t[i]
            delta_sq = delta * delta
            self.m_t[i] = self.beta_1 * m_prev + (1 - self.beta_1) * delta
            self.v_t[i] = self.beta_2 * v_prev + (1 - self.beta_2) * delta_sq
            norm_v_prev = v_prev.abs()
            norm_diff = (delta_sq - self.v_t[i]).abs()
            beta_3 = norm_v_prev / (norm_diff + norm_v_prev)
            beta_3 = torch.nan_to_num(beta_3, nan=0.0)  # 0/0 at the first step (v_prev = 0, delta = 0)
            self.d_t[i] = beta_3 * d_prev + (1 - beta_3) * (delta_sq - self.v_t[i])

    def aggregate_fit(
        self, server_round: int, results: list[tuple[ClientProxy, FitRes]], failures: list[Any]
    ) -> tuple[Parameters | None, dict[str, Scalar]]:
        averaged, metrics = super().aggregate_fit(server_round, results, failures)
        if averaged is None:
            return None, {}
        mean = [to_tensor(w) for w in parameters_to_ndarrays(averaged)]
        if not self.current_weights:
            self._initialise(NDArrays(mean))
        current = [c.to(m.device) for c, m in zip(self.current_weights, mean)]
        self.m_t = [t.to(m.device) for t, m in zip(self.m_t, mean)]
        self.v_t = [t.to(m.device) for t, m in zip(self.v_t, mean)]
        self.d_t = [t.to(m.device) for t, m in zip(self.d_t, mean)]
        self._update_parameters([m.to(torch.float32) - c.to(torch.float32) for m, c in zip(mean, current)])
        new_weights = []
        for c, m, v, d in zip(current, self.m_t, self.v_t, self.d_t):
            if c.is_floating_point():
                new_weights.append((c.to(torch.float32) + self.eta * m / (v.sqrt() - d + self.tau)).to(c.dtype))
            else:
                new_weights.append(c)
        self.current_weights = new_weights
        return ndarrays_to_parameters(NDArrays(new_weights)), metrics
