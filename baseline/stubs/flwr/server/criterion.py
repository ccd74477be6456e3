from __future__ import annotations

from abc import ABC, abstractmethod

from .client_proxy import ClientProxy


class Criterion(ABC):
    """Predicate deciding whether a client is eligible for sampling."""

    @abstractmethod
    def select(self, client: ClientProxy) -> bool: ...
