"""FL-level privacy accountants (parity: ``fl4health/privacy/fl_accountants.py:12-242``).

* ``FlInstanceLevelAccountant``: instance-level DP-SGD inside FL — per client the effective sampling rate is
  ``client_sampling_rate * batch_size / dataset_size``; the reported guarantee is the max over clients.
* ``FlClientLevelAccountantPoissonSampling`` / ``...FixedSamplingNoReplacement``: client-level DP where one "update" is
  one server round; both accept trajectories (lists) of settings.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from math import ceil

from fl4health_b200.privacy.moments_accountant import (
    FixedSamplingWithoutReplacement,
    MomentsAccountant,
    PoissonSampling,
    SamplingStrategy,
)


class FlInstanceLevelAccountant:
    def __init__(
        self,
        client_sampling_rate: float,
        noise_multiplier: float,
        epochs_per_round: int,
        client_batch_sizes: list[int],
        client_dataset_sizes: list[int],
        moment_orders: list[float] | None = None,
    ) -> None:
        assert len(client_batch_sizes) == len(client_dataset_sizes)
        self.noise_multiplier = noise_multiplier
        self.epochs_per_round = epochs_per_round
        self.num_batches_per_client = [ceil(d / b) for b, d in zip(client_batch_sizes, client_dataset_sizes)]
        self.sampling_strategies_per_client = [
            PoissonSampling(client_sampling_rate * b / d) for b, d in zip(client_batch_sizes, client_dataset_sizes)
        ]
        self.accountant = MomentsAccountant(moment_orders)

    def _per_client(self, server_updates: int):  # noqa: ANN202
        for num_batches, strategy in zip(self.num_batches_per_client, self.sampling_strategies_per_client):
            yield strategy, ceil(server_updates * self.epochs_per_round * num_batches)

    def get_epsilon(self, server_updates: int, delta: float) -> float:
        return max(self.accountant.get_epsilon(s, self.noise_multiplier, t, delta) for s, t in self._per_client(server_updates))

    def get_delta(self, server_updates: int, epsilon: float) -> float:
        return max(self.accountant.get_delta(s, self.noise_multiplier, t, epsilon) for s, t in self._per_client(server_updates))


class ClientLevelAccountant(ABC):
    def __init__(self, noise_multiplier: float | list[float], moment_orders: list[float] | None = None) -> None:
        self.noise_multiplier = noise_multiplier
        self.accountant = MomentsAccountant(moment_orders)
        self.sampling_strategy: SamplingStrategy | list

    @abstractmethod
    def get_epsilon(self, server_updates: int | list[int], delta: float) -> float: ...

    @abstractmethod
    def get_delta(self, server_updates: int | list[int], epsilon: float) -> float: ...

    def _validate_server_updates(self, server_updates: int | list[int]) -> None:
        if isinstance(server_updates, list):
            assert isinstance(self.noise_multiplier, list) and len(server_updates) == len(self.noise_multiplier)
        else:
            assert isinstance(self.noise_multiplier, float)


class _ClientLevelAccountantImpl(ClientLevelAccountant):
    def get_epsilon(self, server_updates: int | list[int], delta: float) -> float:
        self._validate_server_updates(server_updates)
        return self.accountant.get_epsilon(self.sampling_strategy, self.noise_multiplier, server_updates, delta)

    def get_delta(self, server_updates: int | list[int], epsilon: float) -> float:
        self._validate_server_updates(server_updates)
        return self.accountant.get_delta(self.sampling_strategy, self.noise_multiplier, server_updates, epsilon)


class FlClientLevelAccountantPoissonSampling(_ClientLevelAccountantImpl):
    def __init__(self, client_sampling_rate: float | list[float], noise_multiplier: float | list[float],
                 moment_orders: list[float] | None = None) -> None:
        super().__init__(noise_multiplier, moment_orders)
        if isinstance(client_sampling_rate, list):
            self.sampling_strategy = [PoissonSampling(q) for q in client_sampling_rate]
        else:
            self.sampling_strategy = PoissonSampling(client_sampling_rate)


class FlClientLevelAccountantFixedSamplingNoReplacement(_ClientLevelAccountantImpl):
    def __init__(self, n_total_clients: int, n_clients_sampled: int | list[int], noise_multiplier: float | list[float],
                 moment_orders: list[float] | None = None) -> None:
        super().__init__(noise_multiplier, moment_orders)
        if isinstance(n_clients_sampled, list):
            self.sampling_strategy = [FixedSamplingWithoutReplacement(n_total_clients, n) for n in n_clients_sampled]
        else:
            self.sampling_strategy = FixedSamplingWithoutReplacement(n_total_clients, n_clients_sampled)
