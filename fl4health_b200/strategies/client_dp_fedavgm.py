"""Client-level DP FedAvg with server momentum and (optionally) adaptive clipping.

Parity: ``fl4health/strategies/client_dp_fedavgm.py:33-467`` (McMahan et al. 2018; Andrew et al. 2021): clients send
clipped weight deltas + clipping bits; the server forms a noised (un)weighted mean, applies momentum
``m <- beta m + Delta`` and ``w <- w + lr m``; with adaptive clipping the bound follows
``C <- C exp(-lr_C (noised_mean_bit - quantile))`` and the weight noise multiplier is corrected to account for the
privacy spent on the bits.  Wire format ``weights ++ [clipping_bound]`` is unchanged.
"""

from __future__ import annotations

import math
from collections.abc import Callable
from logging import INFO
from typing import Any

import numpy as np
import torch

from fl4health_b200.client_managers.base_sampling_manager import BaseFractionSamplingManager
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import (
    EvaluateIns,
    FitIns,
    FitRes,
    MetricsAggregationFn,
    NDArrays,
    Parameters,
    Scalar,
    ndarrays_to_parameters,
    parameters_to_ndarrays,
    to_tensor,
)
from fl4health_b200.parameter_exchange.parameter_packer import ParameterPackerWithClippingBit
from fl4health_b200.servers.client_manager import ClientManager
from fl4health_b200.servers.client_proxy import ClientProxy
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg
from fl4health_b200.strategies.noisy_aggregate import (
    gaussian_noisy_aggregate_clipping_bits,
    gaussian_noisy_unweighted_aggregate,
    gaussian_noisy_weighted_aggregate,
)
from fl4health_b200.utils.functions import decode_and_pseudo_sort_results


class ClientLevelDPFedAvgM(BasicFedAvg):
    def __init__(
        self,
        *,
        fraction_fit: float = 1.0,
        fraction_evaluate: float = 1.0,
        min_available_clients: int = 2,
        evaluate_fn: Callable[[int, NDArrays, dict[str, Scalar]], tuple[float, dict[str, Scalar]] | None] | None = None,
        on_fit_config_fn: Callable[[int], dict[str, Scalar]] | None = None,
        on_evaluate_config_fn: Callable[[int], dict[str, Scalar]] | None = None,
        accept_failures: bool = True,
        initial_parameters: Parameters | None = None,
        fit_metrics_aggregation_fn: MetricsAggregationFn | None = None,
        evaluate_metrics_aggregation_fn: MetricsAggregationFn | None = None,
        weighted_aggregation: bool = False,
        weighted_eval_losses: bool = True,
        per_client_example_cap: float | None = None,
        adaptive_clipping: bool = False,
        server_learning_rate: float = 1.0,
        clipping_learning_rate: float = 1.0,
        clipping_quantile: float = 0.5,
        initial_clipping_bound: float = 0.1,
        weight_noise_multiplier: float = 1.0,
        clipping_noise_multiplier: float = 1.0,
        beta: float = 0.9,
    ) -> None:
        assert 0.0 <= clipping_quantile <= 1.0
        self.clipping_bound = initial_clipping_bound
        self.current_weights: NDArrays = NDArrays()
        if initial_parameters:
            self.add_auxiliary_information(initial_parameters)
        super().__init__(
            fraction_fit=fraction_fit, fraction_evaluate=fraction_evaluate, min_available_clients=min_available_clients,
            evaluate_fn=evaluate_fn, on_fit_config_fn=on_fit_config_fn, on_evaluate_config_fn=on_evaluate_config_fn,
            accept_failures=accept_failures, initial_parameters=initial_parameters,
            fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
            evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn,
            weighted_aggregation=weighted_aggregation, weighted_eval_losses=weighted_eval_losses,
        )
        self.per_client_example_cap = per_client_example_cap
        self.adaptive_clipping = adaptive_clipping
        self.server_learning_rate = server_learning_rate
        self.clipping_learning_rate = clipping_learning_rate
        self.clipping_quantile = clipping_quantile
        self.weight_noise_multiplier = weight_noise_multiplier
        self.clipping_noise_multiplier = clipping_noise_multiplier
        self.beta = beta
        self.parameter_packer = ParameterPackerWithClippingBit()
        self.sample_counts: list[int] | None = None
        self.m_t: NDArrays | None = None
        self.total_client_weight: float = 0.0

    def __repr__(self) -> str:
        return f"ClientLevelDPFedAvgM(accept_failures={self.accept_failures})"

    def add_auxiliary_information(self, original_parameters: Parameters) -> None:
        """Remember the weights as the server's starting point and append the clipping bound for the clients."""
        self.current_weights = NDArrays([to_tensor(w).clone() for w in parameters_to_ndarrays(original_parameters)])
        original_parameters.tensors.append(np.array([self.clipping_bound]))
        original_parameters.flat = None

    def modify_noise_multiplier(self) -> float:
        """z_delta = (z^-2 - (2 z_b)^-2)^-1/2 (Andrew et al. 2021, Thm 1): extra weight noise pays for the noised bits."""
        sqrt_argument = pow(self.weight_noise_multiplier, -2.0) - pow(2.0 * self.clipping_noise_multiplier, -2.0)
        if sqrt_argument < 0.0:
            raise ValueError(
                "Noise Multiplier modification will fail. The relationship of the weight and clipping noise "
                f"multipliers leads to negative sqrt argument {sqrt_argument}"
            )
        return pow(sqrt_argument, -0.5)

    def split_model_weights_and_clipping_bits(
        self, results: list[tuple[ClientProxy, FitRes]]
    ) -> tuple[list[tuple[NDArrays, int]], NDArrays]:
        weights_and_counts: list[tuple[NDArrays, int]] = []
        clipping_bits = NDArrays()
        for _, packed, sample_count in decode_and_pseudo_sort_results(results):
            weights, bit = self.parameter_packer.unpack_parameters(packed)
            weights_and_counts.append((weights, sample_count))
            clipping_bits.append(np.array(bit))
        return weights_and_counts, clipping_bits

    def calculate_update_with_momentum(self, weights_update: NDArrays) -> None:
        if not self.m_t:
            self.m_t = weights_update
        else:
            self.m_t = NDArrays([self.beta * to_tensor(prev) + to_tensor(new, to_tensor(prev).device) for prev, new in zip(self.m_t, weights_update)])

    def update_current_weights(self) -> None:
        assert self.m_t is not None
        updated = NDArrays()
        for current, m in zip(self.current_weights, self.m_t):
            cur = to_tensor(current)
            step = self.server_learning_rate * to_tensor(m, cur.device)
            updated.append((cur.to(torch.float32) + step).to(cur.dtype) if cur.is_floating_point() else cur)
        self.current_weights = updated

    def _update_clipping_bound_with_noised_bits(self, noised_clipping_bits: float) -> None:
        self.clipping_bound *= math.exp(-self.clipping_learning_rate * (noised_clipping_bits - self.clipping_quantile))

    def update_clipping_bound(self, clipping_bits: NDArrays) -> None:
        self._update_clipping_bound_with_noised_bits(
            gaussian_noisy_aggregate_clipping_bits(clipping_bits, self.clipping_noise_multiplier)
        )

    def aggregate_fit(
        self, server_round: int, results: list[tuple[ClientProxy, FitRes]], failures: list[Any]
    ) -> tuple[Parameters | None, dict[str, Scalar]]:
        if not results or (not self.accept_failures and failures):
            return None, {}
        if self.weighted_aggregation and (self.per_client_example_cap is None or self.total_client_weight == 0.0):
            assert self.sample_counts is not None, "poll clients for sample counts before weighted DP aggregation"
            if self.per_client_example_cap is None:
                self.per_client_example_cap = float(sum(self.sample_counts))
            self.total_client_weight = sum(min(n / self.per_client_example_cap, 1.0) for n in self.sample_counts)
        weights_and_counts, clipping_bits = self.split_model_weights_and_clipping_bits(results)
        noise_multiplier = self.weight_noise_multiplier
        if self.adaptive_clipping:
            noise_multiplier = self.modify_noise_multiplier()
            self.update_clipping_bound(clipping_bits)
            log(INFO, f"New Clipping Bound is: {self.clipping_bound}")
        if self.weighted_aggregation:
            assert self.per_client_example_cap is not None
            noised_update = gaussian_noisy_weighted_aggregate(
                weights_and_counts, noise_multiplier, self.clipping_bound, self.fraction_fit,
                self.per_client_example_cap, self.total_client_weight,
            )
        else:
            noised_update = gaussian_noisy_unweighted_aggregate(weights_and_counts, noise_multiplier, self.clipping_bound)
        self.calculate_update_with_momentum(noised_update)
        self.update_current_weights()
        packed = self.parameter_packer.pack_parameters(self.current_weights, self.clipping_bound)
        return ndarrays_to_parameters(packed), self._aggregate_fit_metrics(server_round, results)

    def configure_fit(self, server_round: int, parameters: Parameters, client_manager: ClientManager) -> list[tuple[ClientProxy, FitIns]]:
        assert isinstance(client_manager, BaseFractionSamplingManager)
        config = self.on_fit_config_fn(server_round) if self.on_fit_config_fn is not None else {"current_server_round": server_round}
        fit_ins = FitIns(parameters, config)
        return [(c, fit_ins) for c in client_manager.sample_fraction(self.fraction_fit, self.min_available_clients)]

    def configure_evaluate(self, server_round: int, parameters: Parameters, client_manager: ClientManager) -> list[tuple[ClientProxy, EvaluateIns]]:
        assert isinstance(client_manager, BaseFractionSamplingManager)
        if self.fraction_evaluate == 0.0:
            return []
        config = self.on_evaluate_config_fn(server_round) if self.on_evaluate_config_fn is not None else {"current_server_round": server_round}
        evaluate_ins = EvaluateIns(parameters, config)
        return [(c, evaluate_ins) for c in client_manager.sample_fraction(self.fraction_evaluate, self.min_available_clients)]
