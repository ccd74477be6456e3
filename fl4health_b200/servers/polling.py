"""Property polling fan-out (parity: ``fl4health/servers/polling.py:63-98``; no thread pool needed)."""

from __future__ import annotations

from typing import Any

from fl4health_b200.common.typing import GetPropertiesIns, GetPropertiesRes
from fl4health_b200.servers.client_proxy import ClientProxy
from fl4health_b200.servers.transport import _DEFAULT

PollResultsAndFailures = tuple[list[tuple[ClientProxy, GetPropertiesRes]], list[Any]]


def poll_client(client: ClientProxy, ins: GetPropertiesIns) -> tuple[ClientProxy, GetPropertiesRes]:
    return client, client.get_properties(ins, timeout=None)


def poll_clients(
    client_instructions: list[tuple[ClientProxy, GetPropertiesIns]],
    max_workers: int | None = None,
    timeout: float | None = None,
) -> PollResultsAndFailures:
    return _DEFAULT.poll_clients(client_instructions, max_workers, timeout)
