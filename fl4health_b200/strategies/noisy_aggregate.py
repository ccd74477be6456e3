"""Gaussian-noised aggregation for client-level DP (parity: ``fl4health/strategies/noisy_aggregate.py:20-143``).

``(sum_k c_k Delta_k + N(0, sigma^2 I)) / K``.  The reference draws one full-size ``np.random.normal`` array per layer on
the CPU (SURVEY hot-op C7).  Here the K client updates are folded into ONE flat fp32 vector (one ``weighted_sum`` launch
per call when the layers are CUDA tensors), the noise is added by the counter-based Philox kernel
(``fl4h_add_gaussian``: no noise tensor is ever materialised) and the result is handed back as per-layer views of that
vector.  Seeds (and the scalar noise on the clipping bits) come from the server-side random streams
(``servers/client_manager.py``): seeded with everything else by ``set_all_random_seeds``, identical on every rank of a
replicated SPMD server, and untouched by what clients do with the global generators.
"""

from __future__ import annotations

import torch

from fl4health_b200.common.typing import NDArray, NDArrays, to_tensor
from fl4health_b200.ops import flat as flat_ops
from fl4health_b200.servers.client_manager import sampling_streams



def _next_seed() -> int:
    return sampling_streams.next_kernel_seed()


def _flatten(update: NDArrays, device: torch.device) -> torch.Tensor:
    return torch.cat([to_tensor(layer, device).to(torch.float32).reshape(-1) for layer in update])


def _noised_mean(updates: list[NDArrays], coefficients: list[float], sigma: float, denominator: int) -> NDArrays:
    """``(sum_k coefficients[k] * updates[k] + N(0, sigma^2)) / denominator`` as per-layer views of one flat vector."""
    first = [to_tensor(layer) for layer in updates[0]]
    device = next((t.device for t in first if t.is_cuda), first[0].device)
    shapes = [t.shape for t in first]
    total = sum(t.numel() for t in first)
    padded = (total + 3) // 4 * 4  # the flat kernels work on 16-byte vectors
    flat = torch.zeros(padded, dtype=torch.float32, device=device)
    for update, coefficient in zip(updates, coefficients):
        flat[:total].add_(_flatten(update, device), alpha=coefficient)
    if sigma > 0:
        flat_ops.add_gaussian_(flat, sigma, _next_seed())
    flat.div_(denominator)
    pieces = flat[:total].split([int(torch.Size(shape).numel()) for shape in shapes])
    return NDArrays([piece.view(shape) for piece, shape in zip(pieces, shapes)])


def add_noise_to_array(layer: NDArray, noise_std_dev: float, denominator: int) -> torch.Tensor:
    return _noised_mean([NDArrays([layer])], [1.0], noise_std_dev, denominator)[0]


def add_noise_to_ndarrays(client_model_updates: list[NDArrays], sigma: float, n_clients: int) -> NDArrays:
    return _noised_mean(client_model_updates, [1.0] * len(client_model_updates), sigma, n_clients)


def gaussian_noisy_unweighted_aggregate(results: list[tuple[NDArrays, int]], noise_multiplier: float, clipping_bound: float) -> NDArrays:
    updates = [update for update, _ in results]
    return _noised_mean(updates, [1.0] * len(updates), noise_multiplier * clipping_bound, len(updates))


def gaussian_noisy_weighted_aggregate(
    results: list[tuple[NDArrays, int]], noise_multiplier: float, clipping_bound: float, fraction_fit: float,
    per_client_example_cap: float, total_client_weight: float,
) -> NDArrays:
    """Example-count weights capped at ``per_client_example_cap`` (McMahan et al. 2018, Alg. 1); the sensitivity — hence
    the noise — scales with the largest weight actually present."""
    capped = [min(count / per_client_example_cap, 1.0) for _, count in results]
    normaliser = fraction_fit * total_client_weight
    sigma = noise_multiplier * clipping_bound * max(capped) / fraction_fit
    return _noised_mean([update for update, _ in results], [c / normaliser for c in capped], sigma, len(results))


def gaussian_noisy_aggregate_clipping_bits(bits: NDArrays, noise_std_dev: float) -> float:
    count = len(bits)
    clipped = sum(float(to_tensor(bit).reshape(()).item()) for bit in bits)
    noise = float(torch.randn((), generator=sampling_streams.torch).item()) * noise_std_dev if noise_std_dev > 0 else 0.0
    return (clipped + noise) / count
