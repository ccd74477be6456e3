"""Single-process federation: one server object + K client objects, no ranks, no RPC.

This is the engine's equivalent of the reference's "server process + N client processes over localhost gRPC"
(``examples/utils/run_fl_local.sh``) for CPU runs, tests, and single-GPU simulation of many clients.
"""

from __future__ import annotations

from collections.abc import Sequence
from typing import Any

from fl4health_b200.common.history import History
from fl4health_b200.servers.base_server import FlServer
from fl4health_b200.servers.client_proxy import InProcessClientProxy


def register_clients(server: FlServer, clients: Sequence[Any]) -> list[InProcessClientProxy]:
    proxies = []
    for index, client in enumerate(clients):
        cid = str(getattr(client, "client_name", None) or f"client_{index}")
        proxy = InProcessClientProxy(cid, client)
        server.client_manager().register(proxy)
        proxies.append(proxy)
    return proxies


def run_simulation(
    server: FlServer, clients: Sequence[Any], num_rounds: int, timeout: float | None = None, shutdown: bool = True
) -> History:
    register_clients(server, clients)
    history, _ = server.fit(num_rounds=num_rounds, timeout=timeout)
    if shutdown:
        if hasattr(server, "disconnect_all_clients"):
            server.disconnect_all_clients(timeout=timeout)
        if hasattr(server, "shutdown"):  # evaluation-only servers have no reporters / state to flush
            server.shutdown()
    return history
