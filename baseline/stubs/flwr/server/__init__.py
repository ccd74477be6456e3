from __future__ import annotations

from dataclasses import dataclass
from logging import INFO

from ..common.logger import log
from .client_manager import ClientManager, SimpleClientManager  # noqa: F401
from .history import History  # noqa: F401
from .server import Server  # noqa: F401


@dataclass
class ServerConfig:
    num_rounds: int = 1
    round_timeout: float | None = None


def start_server(*, server_address: str = "0.0.0.0:8080", server: Server | None = None, config: ServerConfig | None = None,  # noqa: S104
                 strategy=None, client_manager: ClientManager | None = None, **_: object) -> History:  # noqa: ANN001
    """Listen, run ``server.fit`` for ``config.num_rounds`` rounds, disconnect every client, return the history."""
    from .._transport import Acceptor

    config = config or ServerConfig()
    if server is None:
        server = Server(client_manager=client_manager or SimpleClientManager(), strategy=strategy)
    acceptor = Acceptor(server_address, server.client_manager())
    acceptor.start()
    log(INFO, "flwr-shim server listening on %s", server_address)
    try:
        history, elapsed = server.fit(num_rounds=config.num_rounds, timeout=config.round_timeout)
        log(INFO, "[SUMMARY] run finished %s round(s) in %.2fs", config.num_rounds, elapsed)
    finally:
        server.disconnect_all_clients(timeout=config.round_timeout)
        acceptor.stop()
    return history
