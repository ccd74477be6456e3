"""Gaussian-noised aggregation for client-level DP (parity: ``fl4health/strategies/noisy_aggregate.py:20-143``).
``(sum_k c_k Delta_k + N(0, sigma^2)) / K`` per layer, computed on the arrays' device (the reference draws full-size
``np.random.normal`` tensors on the CPU, SURVEY hot-op C7)."""

from __future__ import annotations

import torch

from fl4health_b200.common.typing import NDArray, NDArrays, to_tensor


def add_noise_to_array(layer: NDArray, noise_std_dev: float, denominator: int) -> torch.Tensor:
    tensor = to_tensor(layer).to(torch.float32)
    noise = torch.randn(tensor.shape, device=tensor.device, dtype=tensor.dtype) * noise_std_dev if noise_std_dev > 0 else 0.0
    return (tensor + noise) / denominator


def add_noise_to_ndarrays(client_model_updates: list[NDArrays], sigma: float, n_clients: int) -> NDArrays:
    out = NDArrays()
    for layer_updates in zip(*client_model_updates):
        tensors = [to_tensor(u).to(torch.float32) for u in layer_updates]
        total = tensors[0].clone()
        for t in tensors[1:]:
            total = total + t.to(total.device)
        out.append(add_noise_to_array(total, sigma, n_clients))
    return out


def gaussian_noisy_unweighted_aggregate(results: list[tuple[NDArrays, int]], noise_multiplier: float, clipping_bound: float) -> NDArrays:
    return add_noise_to_ndarrays([nds for nds, _ in results], noise_multiplier * clipping_bound, len(results))


def gaussian_noisy_weighted_aggregate(
    results: list[tuple[NDArrays, int]], noise_multiplier: float, clipping_bound: float, fraction_fit: float,
    per_client_example_cap: float, total_client_weight: float,
) -> NDArrays:
    coefficients = [min(n_points / per_client_example_cap, 1.0) for _, n_points in results]
    scaled = [c / (fraction_fit * total_client_weight) for c in coefficients]
    updates = [NDArrays([to_tensor(layer) * coef for layer in nds]) for (nds, _), coef in zip(results, scaled)]
    sigma = (noise_multiplier * clipping_bound * max(coefficients)) / fraction_fit
    return add_noise_to_ndarrays(updates, sigma, len(results))


def gaussian_noisy_aggregate_clipping_bits(bits: NDArrays, noise_std_dev: float) -> float:
    total = sum(float(to_tensor(b).reshape(()).item()) for b in bits)
    return float(add_noise_to_array(torch.tensor(total), noise_std_dev, len(bits)).item())
