"""Strategies that can poll all clients for properties (parity: ``strategy_with_poll.py:8-18``)."""

from __future__ import annotations

from abc import ABC, abstractmethod

from fl4health_b200.common.typing import GetPropertiesIns
from fl4health_b200.servers.client_manager import ClientManager
from fl4health_b200.servers.client_proxy import ClientProxy


class StrategyWithPolling(ABC):
    @abstractmethod
    def configure_poll(
        self, server_round: int, client_manager: ClientManager
    ) -> list[tuple[ClientProxy, GetPropertiesIns]]:
        raise NotImplementedError
