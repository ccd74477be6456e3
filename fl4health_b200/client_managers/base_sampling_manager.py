"""Fraction-based sampling managers (parity: ``fl4health/client_managers/base_sampling_manager.py:8-87``)."""

from __future__ import annotations

from logging import INFO

from fl4health_b200.common.logger import log
from fl4health_b200.servers.client_manager import Criterion, SimpleClientManager
from fl4health_b200.servers.client_proxy import ClientProxy


class BaseFractionSamplingManager(SimpleClientManager):
    """Managers that sample a *fraction* of clients; the count-based ``sample`` is deliberately unavailable."""

    def sample(self, num_clients: int, min_num_clients: int | None = None, criterion: Criterion | None = None) -> list[ClientProxy]:
        raise NotImplementedError(
            "The basic sampling function is not implemented for these managers. Please use the fraction sample function"
        )

    def wait_and_filter(self, min_num_clients: int | None, criterion: Criterion | None = None) -> list[str]:
        if min_num_clients is not None:
            self.wait_for(min_num_clients)
        else:
            self.wait_for(1)
        available_cids = sorted(self.clients)  # stable order: every SPMD rank must draw the same sample
        if criterion is not None:
            available_cids = [cid for cid in available_cids if criterion.select(self.clients[cid])]
        return available_cids

    def sample_one(self, min_num_clients: int | None = None, criterion: Criterion | None = None) -> list[ClientProxy]:
        return super().sample(1, min_num_clients, criterion)

    def sample_fraction(
        self, sample_fraction: float, min_num_clients: int | None = None, criterion: Criterion | None = None
    ) -> list[ClientProxy]:
        raise NotImplementedError

    def sample_all(self, min_num_clients: int | None = None, criterion: Criterion | None = None) -> list[ClientProxy]:
        available_cids = self.wait_and_filter(min_num_clients, criterion)
        if not available_cids:
            log(INFO, "No clients available for sampling")
        return [self.clients[cid] for cid in available_cids]
