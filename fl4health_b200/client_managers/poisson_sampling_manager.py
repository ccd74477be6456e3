"""Poisson (iid Bernoulli) client sampling (parity: ``poisson_sampling_manager.py:11-50``).  May select 0..N clients."""

from __future__ import annotations

from logging import INFO

from fl4health_b200.client_managers.base_sampling_manager import BaseFractionSamplingManager
from fl4health_b200.common.logger import log
from fl4health_b200.servers.client_manager import Criterion, sampling_streams
from fl4health_b200.servers.client_proxy import ClientProxy


class PoissonSamplingClientManager(BaseFractionSamplingManager):
    def _poisson_sample(self, sampling_probability: float, available_cids: list[str]) -> list[str]:
        draws = sampling_streams.numpy.binomial(1, sampling_probability, len(available_cids)).astype(bool)
        return [cid for cid, keep in zip(available_cids, draws) if keep]

    def sample_fraction(
        self, sample_fraction: float, min_num_clients: int | None = None, criterion: Criterion | None = None
    ) -> list[ClientProxy]:
        available_cids = self.wait_and_filter(min_num_clients, criterion)
        if not available_cids:
            return []
        sampled = self._poisson_sample(sample_fraction, available_cids)
        if not sampled:
            log(INFO, f"Sampling was successful but no clients were selected: probability {sample_fraction}")
        return [self.clients[cid] for cid in sampled]
