"""Fixed-size sampling without replacement: floor(q*N) clients (parity: ``fixed_without_replacement_manager.py:11-48``)."""

from __future__ import annotations

import math
from logging import INFO

from fl4health_b200.client_managers.base_sampling_manager import BaseFractionSamplingManager
from fl4health_b200.common.logger import log
from fl4health_b200.servers.client_manager import Criterion, sampling_streams
from fl4health_b200.servers.client_proxy import ClientProxy


class FixedSamplingByFractionClientManager(BaseFractionSamplingManager):
    def sample_fraction(
        self, sample_fraction: float, min_num_clients: int | None = None, criterion: Criterion | None = None
    ) -> list[ClientProxy]:
        available_cids = self.wait_and_filter(min_num_clients, criterion)
        if not available_cids:
            return []
        n_clients = math.floor(len(available_cids) * sample_fraction)
        if n_clients == 0:
            log(INFO, f"Sample fraction {sample_fraction} of {len(available_cids)} clients selects no one.")
            return []
        return [self.clients[cid] for cid in sampling_streams.python.sample(available_cids, n_clients)]
