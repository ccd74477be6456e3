"""Moments (RDP) accountant over sampling strategies.

API parity: ``fl4health/privacy/moments_accountant.py:64-258`` (``PoissonSampling``,
``FixedSamplingWithoutReplacement``, ``MomentsAccountant.get_epsilon/get_delta`` for a single setting or a trajectory
of settings, default orders 1.25 ... 512).  The RDP math lives in ``privacy/rdp.py``.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from collections.abc import Sequence

import numpy as np

from fl4health_b200.privacy import rdp
from fl4health_b200.privacy.dp_events import (
    DpEvent,
    GaussianDpEvent,
    NeighborRel,
    PoissonSampledDpEvent,
    SampledWithoutReplacementDpEvent,
    SelfComposedDpEvent,
)


class SamplingStrategy(ABC):
    neighbor_relation: NeighborRel

    @abstractmethod
    def rdp_per_update(self, noise_multiplier: float, orders: Sequence[float]) -> np.ndarray:
        raise NotImplementedError

    @abstractmethod
    def get_dp_event(self, noise_event: DpEvent) -> DpEvent:
        """The sampled mechanism as a plain-data event (parity: ``moments_accountant.py:26-61``)."""
        raise NotImplementedError

    def composed_event(self, noise_multiplier: float, updates: int) -> DpEvent:
        """``updates`` self-compositions of the sampled Gaussian mechanism: what the accountant charges for."""
        return SelfComposedDpEvent(self.get_dp_event(GaussianDpEvent(noise_multiplier)), updates)


class PoissonSampling(SamplingStrategy):
    """Each element participates independently with probability ``sampling_ratio`` (add/remove-one neighbours)."""

    neighbor_relation = NeighborRel.ADD_OR_REMOVE_ONE

    def __init__(self, sampling_ratio: float) -> None:
        self.sampling_ratio = sampling_ratio

    def get_dp_event(self, noise_event: DpEvent) -> DpEvent:
        return PoissonSampledDpEvent(self.sampling_ratio, noise_event)

    def rdp_per_update(self, noise_multiplier: float, orders: Sequence[float]) -> np.ndarray:
        return rdp.rdp_poisson_subsampled_gaussian(self.sampling_ratio, noise_multiplier, orders)


class FixedSamplingWithoutReplacement(SamplingStrategy):
    """Exactly ``sample_size`` of ``population_size`` elements per update (replace-one neighbours)."""

    neighbor_relation = NeighborRel.REPLACE_ONE

    def __init__(self, population_size: int, sample_size: int) -> None:
        self.population_size = population_size
        self.sample_size = sample_size

    def get_dp_event(self, noise_event: DpEvent) -> DpEvent:
        return SampledWithoutReplacementDpEvent(self.population_size, self.sample_size, noise_event)

    def rdp_per_update(self, noise_multiplier: float, orders: Sequence[float]) -> np.ndarray:
        return rdp.rdp_sample_wor_gaussian(self.sample_size / self.population_size, noise_multiplier, orders)


class MomentsAccountant:
    def __init__(self, moment_orders: list[float] | None = None) -> None:
        if moment_orders is not None:
            self.moment_orders = moment_orders
        else:
            low = [1.25, 1.5, 1.75, 2.0, 2.25, 2.5, 3.0, 3.5, 4.0, 4.5]
            self.moment_orders = low + [float(x) for x in range(5, 64)] + [128.0, 256.0, 512.0]

    def _validate_accountant_input(self, sampling_strategies, noise_multiplier, updates) -> None:  # noqa: ANN001
        all_lists = isinstance(sampling_strategies, Sequence) and isinstance(noise_multiplier, list) and isinstance(updates, list)
        all_values = isinstance(sampling_strategies, SamplingStrategy) and isinstance(noise_multiplier, float) and isinstance(updates, int)
        assert all_lists or all_values, "pass either single values or equal-length lists (a trajectory)"

    def _total_rdp(self, sampling_strategies, noise_multipliers, updates) -> np.ndarray:  # noqa: ANN001
        if isinstance(sampling_strategies, SamplingStrategy):
            sampling_strategies, noise_multipliers, updates = [sampling_strategies], [noise_multipliers], [updates]
        total = np.zeros(len(self.moment_orders))
        for strategy, z, t in zip(sampling_strategies, noise_multipliers, updates):
            total = total + t * strategy.rdp_per_update(float(z), self.moment_orders)  # self-composition is additive
        return total

    def get_epsilon(self, sampling_strategies, noise_multiplier, updates, delta: float) -> float:  # noqa: ANN001
        self._validate_accountant_input(sampling_strategies, noise_multiplier, updates)
        return rdp.epsilon_from_rdp(self.moment_orders, self._total_rdp(sampling_strategies, noise_multiplier, updates), delta)

    def get_delta(self, sampling_strategies, noise_multiplier, updates, epsilon: float) -> float:  # noqa: ANN001
        self._validate_accountant_input(sampling_strategies, noise_multiplier, updates)
        return rdp.delta_from_rdp(self.moment_orders, self._total_rdp(sampling_strategies, noise_multiplier, updates), epsilon)
