"""Localhost TCP stand-in for Flower's gRPC bidi stream.

Server side: an accept thread registers one :class:`SocketClientProxy` per connecting client with the server's
``ClientManager``.  Client side: ``start_client`` connects, then serves ``(verb, ins)`` requests until told to
disconnect.  Frames are ``multiprocessing.connection`` messages (4-byte length prefix + pickle of the dataclass; the
parameter payload inside is already ``np.save`` bytes, so pickling is a memcpy of those blobs).
"""

from __future__ import annotations

import threading
import time
import uuid
from multiprocessing.connection import Client as _Connect
from multiprocessing import AuthenticationError
from multiprocessing.connection import Connection, Listener

from .common.typing import (
    DisconnectRes,
    EvaluateIns,
    EvaluateRes,
    FitIns,
    FitRes,
    GetParametersIns,
    GetParametersRes,
    GetPropertiesIns,
    GetPropertiesRes,
    ReconnectIns,
)
from .server.client_proxy import ClientProxy

_AUTH = b"flwr-shim"


def _split(address: str) -> tuple[str, int]:
    host, _, port = address.rpartition(":")
    host = host.strip("[]") or "127.0.0.1"
    if host == "0.0.0.0":  # noqa: S104
        host = "127.0.0.1"
    return host, int(port)


class SocketClientProxy(ClientProxy):
    """Server-side handle of one connected client; one request in flight at a time (as with the gRPC bridge)."""

    def __init__(self, cid: str, conn: Connection) -> None:
        super().__init__(cid)
        self.conn = conn
        self.lock = threading.Lock()

    def _call(self, verb: str, ins: object, timeout: float | None):  # noqa: ANN202
        with self.lock:
            self.conn.send((verb, ins))
            if timeout is not None and not self.conn.poll(timeout):
                raise TimeoutError(f"client {self.cid} did not answer {verb} within {timeout}s")
            ok, payload = self.conn.recv()
        if not ok:
            raise RuntimeError(f"client {self.cid} raised during {verb}: {payload}")
        return payload

    def get_properties(self, ins: GetPropertiesIns, timeout: float | None, group_id: int | None) -> GetPropertiesRes:
        return self._call("get_properties", ins, timeout)

    def get_parameters(self, ins: GetParametersIns, timeout: float | None, group_id: int | None) -> GetParametersRes:
        return self._call("get_parameters", ins, timeout)

    def fit(self, ins: FitIns, timeout: float | None, group_id: int | None) -> FitRes:
        return self._call("fit", ins, timeout)

    def evaluate(self, ins: EvaluateIns, timeout: float | None, group_id: int | None) -> EvaluateRes:
        return self._call("evaluate", ins, timeout)

    def reconnect(self, ins: ReconnectIns, timeout: float | None, group_id: int | None) -> DisconnectRes:
        with self.lock:
            try:
                self.conn.send(("disconnect", ins))
                self.conn.close()
            except OSError:
                pass
        return DisconnectRes(reason="RECONNECT" if ins.seconds else "ACK")


class Acceptor:
    """Accept loop of the server process."""

    def __init__(self, address: str, client_manager) -> None:  # noqa: ANN001
        # a real backlog: every client of the federation connects within the same few milliseconds (the default of 1
        # makes the kernel drop SYNs; clients then retry while their abandoned half-open attempts are still queued)
        self.listener = Listener(_split(address), family="AF_INET", backlog=256, authkey=_AUTH)
        self.client_manager = client_manager
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.stopping = False

    def start(self) -> None:
        self.thread.start()

    def _loop(self) -> None:
        while not self.stopping:
            try:
                conn = self.listener.accept()  # includes the authentication handshake
                hello = conn.recv()
            except (EOFError, AuthenticationError, ConnectionError):
                continue  # an attempt its client gave up on (it retries on a new connection): not a reason to stop accepting
            except OSError:
                if self.stopping:
                    return
                continue
            cid = str(hello.get("cid") or uuid.uuid4().hex)
            self.client_manager.register(SocketClientProxy(cid, conn))

    def stop(self) -> None:
        self.stopping = True
        try:
            self.listener.close()
        except OSError:
            pass


def start_client(*, server_address: str, client, cid: str | None = None, connect_retries: int = 600, **_: object) -> None:  # noqa: ANN001
    """Connect to the server and serve requests until it disconnects us."""
    if hasattr(client, "to_client"):
        client = client.to_client()
    conn = None
    for _attempt in range(connect_retries):
        try:
            conn = _Connect(_split(server_address), family="AF_INET", authkey=_AUTH)
            break
        except (ConnectionRefusedError, OSError, EOFError, AuthenticationError):
            time.sleep(0.1)
    if conn is None:
        raise ConnectionError(f"could not reach flwr-shim server at {server_address}")
    conn.send({"cid": cid})
    while True:
        try:
            verb, ins = conn.recv()
        except (EOFError, OSError):
            break
        if verb == "disconnect":
            break
        try:
            conn.send((True, getattr(client, verb)(ins)))
        except Exception as exc:  # noqa: BLE001
            import traceback

            conn.send((False, "".join(traceback.format_exception(exc))))
    shutdown = getattr(getattr(client, "numpy_client", None), "shutdown", None)
    if callable(shutdown):
        shutdown()
    try:
        conn.close()
    except OSError:
        pass
