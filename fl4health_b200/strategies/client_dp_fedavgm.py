"""Client-level DP FedAvg with server momentum and (optionally) adaptive clipping.

Parity: ``fl4health/strategies/client_dp_fedavgm.py:33-467`` (McMahan et al. 2018; Andrew et al. 2021): clients send
clipped weight deltas + clipping bits; the server forms a noised (un)weighted mean, applies momentum
``m <- beta m + Delta`` and ``w <- w + lr m``; with adaptive clipping the bound follows
``C <- C exp(-lr_C (noised_mean_bit - quantile))`` and the weight noise multiplier is corrected to account for the
privacy spent on the bits.  Wire format ``weights ++ [clipping_bound]`` is unchanged.
"""

from __future__ import annotations

import math
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
        initial_parameters: Parameters | None = None,
        weighted_aggregation: bool = False,
        per_client_example_cap: float | None = None,
        adaptive_clipping: bool = False,
        server_learning_rate: float = 1.0,
        clipping_learning_rate: float = 1.0,
        clipping_quantile: float = 0.5,
        initial_clipping_bound: float = 0.1,
        weight_noise_multiplier: float = 1.0,
        clipping_noise_multiplier: float = 1.0,
        beta: float = 0.9,
        **fedavg_options: Any,
    ) -> None:
        """``fedavg_options``: the remaining ``BasicFedAvg`` keywords (``fraction_fit``, ``fraction_evaluate``,
        ``min_available_clients``, ``evaluate_fn``, config / metric-aggregation functions, ``weighted_eval_losses`` ...).
        Sampling is by fraction through a ``BaseFractionSamplingManager``: ``min_fit_clients`` does not apply."""
        if not 0.0 <= clipping_quantile <= 1.0:
            raise AssertionError("clipping_quantile must lie in [0, 1]")
        self.clipping_bound = initial_clipping_bound
        self.current_weights: NDArrays = NDArrays()
        if initial_parameters:
            self.add_auxiliary_information(initial_parameters)
        super().__init__(initial_parameters=initial_parameters, weighted_aggregation=weighted_aggregation, **fedavg_options)
        self.per_client_example_cap, self.adaptive_clipping = per_client_example_cap, adaptive_clipping
        self.server_learning_rate, self.beta = server_learning_rate, beta
        self.clipping_learning_rate, self.clipping_quantile = clipping_learning_rate, clipping_quantile
        self.weight_noise_multiplier, self.clipping_noise_multiplier = weight_noise_multiplier, clipping_noise_multiplier
        self.parameter_packer = ParameterPackerWithClippingBit()
        self.sample_counts: list[int] | None = None
        self.m_t: NDArrays | None = None
        self.total_client_weight: float = 0.0

    def __repr__(self) -> str:
        return f"ClientLevelDPFedAvgM(accept_failures={self.accept_failures})"

    # ------------------------------------------------------------------------------------------ wire format
    def add_auxiliary_information(self, original_parameters: Parameters) -> None:
        """Remember the weights as the server's starting point and append the clipping bound for the clients."""
        self.current_weights = NDArrays([to_tensor(w).clone() for w in parameters_to_ndarrays(original_parameters)])
        original_parameters.tensors.append(np.array([self.clipping_bound]))
        original_parameters.flat = None

    def split_model_weights_and_clipping_bits(
        self, results: list[tuple[ClientProxy, FitRes]]
    ) -> tuple[list[tuple[NDArrays, int]], NDArrays]:
        deltas: list[tuple[NDArrays, int]] = []
        bits = NDArrays()
        for _, packed, count in decode_and_pseudo_sort_results(results):
            delta, bit = self.parameter_packer.unpack_parameters(packed)
            deltas.append((delta, count))
            bits.append(np.array(bit))
        return deltas, bits

    # ------------------------------------------------------------------------------------------ privacy knobs
    def modify_noise_multiplier(self) -> float:
        """z_delta = (z^-2 - (2 z_b)^-2)^-1/2 (Andrew et al. 2021, Thm 1): extra weight noise pays for the noised bits."""
        radicand = self.weight_noise_multiplier ** -2.0 - (2.0 * self.clipping_noise_multiplier) ** -2.0
        if radicand < 0.0:
            raise ValueError("Noise Multiplier modification will fail. The relationship of the weight and clipping noise "
                             f"multipliers leads to negative sqrt argument {radicand}")
        return radicand ** -0.5

    def _update_clipping_bound_with_noised_bits(self, noised_clipping_bits: float) -> None:
        """Geometric update towards the target quantile: C <- C exp(-eta_C (b_noised - gamma))."""
        self.clipping_bound *= math.exp(-self.clipping_learning_rate * (noised_clipping_bits - self.clipping_quantile))

    def update_clipping_bound(self, clipping_bits: NDArrays) -> None:
        noised = gaussian_noisy_aggregate_clipping_bits(clipping_bits, self.clipping_noise_multiplier)
        self._update_clipping_bound_with_noised_bits(noised)

    def _resolve_client_weights(self) -> None:
        """Weighted aggregation: the example cap defaults to the federation's total sample count; W = sum of capped weights."""
        assert self.sample_counts is not None, "poll clients for sample counts before weighted DP aggregation"
        if self.per_client_example_cap is None:
            self.per_client_example_cap = float(sum(self.sample_counts))
        self.total_client_weight = sum(min(n / self.per_client_example_cap, 1.0) for n in self.sample_counts)

    # ------------------------------------------------------------------------------------------ server optimizer
    def calculate_update_with_momentum(self, weights_update: NDArrays) -> None:
        """m <- beta m + Delta (first round: m = Delta)."""
        if not self.m_t:
            self.m_t = weights_update
            return
        blended = NDArrays()
        for previous, fresh in zip(self.m_t, weights_update):
            previous = to_tensor(previous)
            blended.append(torch.add(to_tensor(fresh, previous.device), previous, alpha=self.beta))
        self.m_t = blended

    def update_current_weights(self) -> None:
        """w <- w + eta m on the floating-point entries (integer buffers are carried over unchanged)."""
        assert self.m_t is not None
        stepped = NDArrays()
        for weights, momentum in zip(self.current_weights, self.m_t):
            weights = to_tensor(weights)
            if weights.is_floating_point():
                moved = torch.add(weights.to(torch.float32), to_tensor(momentum, weights.device), alpha=self.server_learning_rate)
                weights = moved.to(weights.dtype)
            stepped.append(weights)
        self.current_weights = stepped

    def aggregate_fit(
        self, server_round: int, results: list[tuple[ClientProxy, FitRes]], failures: list[Any]
    ) -> tuple[Parameters | None, dict[str, Scalar]]:
        if not results or (failures and not self.accept_failures):
            return None, {}
        if self.weighted_aggregation and (self.per_client_example_cap is None or self.total_client_weight == 0.0):
            self._resolve_client_weights()
        deltas, bits = self.split_model_weights_and_clipping_bits(results)
        noise_multiplier = self.weight_noise_multiplier
        if self.adaptive_clipping:
            noise_multiplier = self.modify_noise_multiplier()
            self.update_clipping_bound(bits)
            log(INFO, f"New Clipping Bound is: {self.clipping_bound}")
        if self.weighted_aggregation:
            assert self.per_client_example_cap is not None
            noised = gaussian_noisy_weighted_aggregate(deltas, noise_multiplier, self.clipping_bound, self.fraction_fit,
                                                       self.per_client_example_cap, self.total_client_weight)
        else:
            noised = gaussian_noisy_unweighted_aggregate(deltas, noise_multiplier, self.clipping_bound)
        self.calculate_update_with_momentum(noised)
        self.update_current_weights()
        outgoing = self.parameter_packer.pack_parameters(self.current_weights, self.clipping_bound)
        return ndarrays_to_parameters(outgoing), self._aggregate_fit_metrics(server_round, results)

    # ------------------------------------------------------------------------------------------ sampling
    def _sampled(self, client_manager: ClientManager, fraction: float, config_fn: Any, server_round: int,
                 parameters: Parameters, instruction: type) -> list:
        assert isinstance(client_manager, BaseFractionSamplingManager)
        config = config_fn(server_round) if config_fn is not None else {"current_server_round": server_round}
        order = instruction(parameters, config)
        return [(proxy, order) for proxy in client_manager.sample_fraction(fraction, self.min_available_clients)]

    def configure_fit(self, server_round: int, parameters: Parameters, client_manager: ClientManager) -> list[tuple[ClientProxy, FitIns]]:
        return self._sampled(client_manager, self.fraction_fit, self.on_fit_config_fn, server_round, parameters, FitIns)

    def configure_evaluate(self, server_round: int, parameters: Parameters, client_manager: ClientManager) -> list[tuple[ClientProxy, EvaluateIns]]:
        assert isinstance(client_manager, BaseFractionSamplingManager)
        if self.fraction_evaluate == 0.0:
            return []
        return self._sampled(client_manager, self.fraction_evaluate, self.on_evaluate_config_fn, server_round, parameters, EvaluateIns)
