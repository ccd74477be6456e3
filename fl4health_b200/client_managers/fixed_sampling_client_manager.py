"""Sample once, reuse until ``reset_sample()`` so fit and evaluate hit the same clients (FedDG-GA)
(parity: ``fixed_sampling_client_manager.py:6-42``)."""

from __future__ import annotations

from fl4health_b200.servers.client_manager import Criterion, SimpleClientManager
from fl4health_b200.servers.client_proxy import ClientProxy


class FixedSamplingClientManager(SimpleClientManager):
    def __init__(self) -> None:
        super().__init__()
        self.current_sample: list[ClientProxy] | None = None

    def reset_sample(self) -> None:
        self.current_sample = None

    def sample(
        self, num_clients: int, min_num_clients: int | None = None, criterion: Criterion | None = None
    ) -> list[ClientProxy]:
        if self.current_sample is None:
            self.current_sample = super().sample(num_clients, min_num_clients, criterion)
        return self.current_sample
