"""Client registry + sampling (role of ``flwr.server.client_manager``)."""

from __future__ import annotations

import random
import threading
from abc import ABC, abstractmethod
from logging import INFO

import numpy as np

from fl4health_b200.common.logger import log
from fl4health_b200.servers.client_proxy import ClientProxy


class _SamplingStreams:
    """The random streams client sampling draws from -- its own, not the process-global ``random`` / ``np.random``.

    With the server logic replicated on every rank (``parallel/spmd.py``) all ranks must draw the same cohort each
    round.  Global generators cannot promise that: anything rank-local that consumes them between rounds (a client's
    data loaders, a Dirichlet partitioner, user augmentation) desynchronises the ranks, which then disagree on who
    trains and mismatch collectives.  ``seed()`` is called with the broadcast seed when a federation is built and by
    ``set_all_random_seeds``; unseeded, the streams start from one draw of the global ``random`` (so a plain
    ``random.seed(...)`` before building the server still fixes the cohorts)."""

    def __init__(self) -> None:
        self._python: random.Random | None = None
        self._numpy: np.random.Generator | None = None

    def seed(self, seed: int | None) -> None:
        if seed is None:
            self._python = self._numpy = None
        else:
            self._python, self._numpy = random.Random(seed), np.random.default_rng(seed)

    def _ensure(self) -> None:
        if self._python is None or self._numpy is None:
            self.seed(random.getrandbits(63))

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
