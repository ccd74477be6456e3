"""Pool of connected clients + uniform sampling."""

from __future__ import annotations

import random
import threading
from abc import ABC, abstractmethod
from logging import INFO

from ..common.logger import log
from .client_proxy import ClientProxy
from .criterion import Criterion


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
    def wait_for(self, num_clients: int, timeout: int) -> bool: ...

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

    def sample(self, num_clients: int, min_num_clients: int | None = None, criterion: Criterion | None = None) -> list[ClientProxy]:
        self.wait_for(num_clients if min_num_clients is None else min_num_clients)
        cids = list(self.clients)
        if criterion is not None:
            cids = [cid for cid in cids if criterion.select(self.clients[cid])]
        if num_clients > len(cids):
            log(INFO, "Sampling failed: number of available clients (%s) is less than number of requested clients (%s).",
                len(cids), num_clients)
            return []
        return [self.clients[cid] for cid in random.sample(cids, num_clients)]
