"""Client registry + sampling (role of ``flwr.server.client_manager``)."""

from __future__ import annotations

import random
import threading
from abc import ABC, abstractmethod
from typing import Any
from logging import INFO

import numpy as np

from fl4health_b200.common.logger import log
from fl4health_b200.servers.client_proxy import ClientProxy


class _SamplingStreams:
    """The random streams SERVER-side logic draws from (client sampling, DP noise on aggregates) -- its own, not the
    process-global ``random`` / ``np.random`` / torch generators.

    With the server logic replicated on every rank (``parallel/spmd.py``) all ranks must draw the same cohort each
    round.  Global generators cannot promise that: anything rank-local that consumes them between rounds (a client's
    data loaders, a Dirichlet partitioner, user augmentation) desynchronises the ranks, which then disagree on who
    trains and mismatch collectives.  ``seed()`` is called with the broadcast seed when a federation is built and by
    ``set_all_random_seeds``; unseeded, the streams start from one draw of the global ``random`` (so a plain
    ``random.seed(...)`` before building the server still fixes the cohorts)."""

    def __init__(self) -> None:
        self._python: random.Random | None = None
        self._numpy: np.random.Generator | None = None
        self._torch: Any = None
        self._seed = 0
        self._draws = 0

    def seed(self, seed: int | None) -> None:
        if seed is None:
            self._python = self._numpy = self._torch = None
        else:
            import torch

            self._python, self._numpy = random.Random(seed), np.random.default_rng(seed)
            self._torch = torch.Generator().manual_seed(seed)
            self._seed, self._draws = int(seed), 0

    def _ensure(self) -> None:
        if self._python is None or self._numpy is None:
            self.seed(random.getrandbits(63))

    @property
    def base_seed(self) -> int:
        """The seed the streams were (or, unseeded, are now) started from: what rank 0 shares with the other ranks."""
        self._ensure()
        return self._seed

    @property
    def python(self) -> random.Random:
        self._ensure()
        assert self._python is not None
        return self._python

    @property
    def numpy(self) -> np.random.Generator:
        self._ensure()
        assert self._numpy is not None
        return self._numpy

    @property
    def torch(self) -> Any:
        """CPU ``torch.Generator`` for small server-side draws (identical on every rank of a replicated server)."""
        self._ensure()
        return self._torch

    def next_kernel_seed(self) -> int:
        """A fresh 63-bit seed for a counter-based (Philox) kernel: same sequence on every rank, new value per call."""
        self._ensure()
        self._draws += 1
        return (self._seed * 1_000_003 + self._draws) & (2**63 - 1)


sampling_streams = _SamplingStreams()


class Criterion(ABC):
    @abstractmethod
    def select(self, client: ClientProxy) -> bool:
        raise NotImplementedError


class ClientManager(ABC):
    @abstractmethod
    def num_available(self) -> int: ...

    @abstractmethod
    def register(self, client: ClientProxy) -> bool: ...

    @abstractmethod
    def unregister(self, client: ClientProxy) -> None: ...

    @abstractmethod
    def all(self) -> dict[str, ClientProxy]: ...

    @abstractmethod
    def wait_for(self, num_clients: int, timeout: int = 86400) -> bool: ...

    @abstractmethod
    def sample(self, num_clients: int, min_num_clients: int | None = None, criterion: Criterion | None = None) -> list[ClientProxy]: ...


class SimpleClientManager(ClientManager):
    def __init__(self) -> None:
        self.clients: dict[str, ClientProxy] = {}
        self._cv = threading.Condition()

    def __len__(self) -> int:
        return len(self.clients)

    def num_available(self) -> int:
        return len(self)

    def wait_for(self, num_clients: int, timeout: int = 86400) -> bool:
        with self._cv:
            return self._cv.wait_for(lambda: len(self.clients) >= num_clients, timeout=timeout)

    def register(self, client: ClientProxy) -> bool:
        if client.cid in self.clients:
            return False
        self.clients[client.cid] = client
        with self._cv:
            self._cv.notify_all()
        return True

    def unregister(self, client: ClientProxy) -> None:
        if client.cid in self.clients:
            del self.clients[client.cid]
            with self._cv:
                self._cv.notify_all()

    def all(self) -> dict[str, ClientProxy]:
        return self.clients

    def sample(
        self, num_clients: int, min_num_clients: int | None = None, criterion: Criterion | None = None
    ) -> list[ClientProxy]:
        if min_num_clients is None:
            min_num_clients = num_clients
        self.wait_for(min_num_clients, timeout=0 if len(self.clients) >= min_num_clients else 86400)
        available = sorted(self.clients)  # stable ids -> deterministic under a fixed seed on every rank
        if criterion is not None:
            available = [cid for cid in available if criterion.select(self.clients[cid])]
        if num_clients > len(available):
            log(INFO, f"Sampling failed: available clients ({len(available)}) < requested clients ({num_clients}).")
            return []
        return [self.clients[cid] for cid in sampling_streams.python.sample(available, num_clients)]
