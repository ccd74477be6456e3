"""Server-side handles for clients (role of ``flwr.server.client_proxy.ClientProxy``).

There is no RPC: ``InProcessClientProxy`` calls the client object directly and hands over live tensors; the SPMD
runtime (``fl4health_b200.parallel.spmd``) provides a rank-addressed proxy with the same interface.  Client ids are
*stable* (client name / rank), unlike Flower's random UUIDs, so per-client strategy state survives restarts.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any

from fl4health_b200.common.typing import (
    Code,
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
    Status,
    ndarrays_to_parameters,
    parameters_to_ndarrays,
)


class ClientProxy(ABC):
    def __init__(self, cid: str) -> None:
        self.cid = cid
        self.properties: dict[str, Any] = {}

    @abstractmethod
    def get_properties(self, ins: GetPropertiesIns, timeout: float | None = None, group_id: int | None = None) -> GetPropertiesRes:
        raise NotImplementedError

    @abstractmethod
    def get_parameters(self, ins: GetParametersIns, timeout: float | None = None, group_id: int | None = None) -> GetParametersRes:
        raise NotImplementedError

    @abstractmethod
    def fit(self, ins: FitIns, timeout: float | None = None, group_id: int | None = None) -> FitRes:
        raise NotImplementedError

    @abstractmethod
    def evaluate(self, ins: EvaluateIns, timeout: float | None = None, group_id: int | None = None) -> EvaluateRes:
        raise NotImplementedError

    def reconnect(self, ins: ReconnectIns, timeout: float | None = None, group_id: int | None = None) -> DisconnectRes:
        return DisconnectRes(reason="")


class InProcessClientProxy(ClientProxy):
    """Direct, zero-copy calls into a client living in this process."""

    def __init__(self, cid: str, client: Any) -> None:
        super().__init__(cid)
        self.client = client

    def get_properties(self, ins: GetPropertiesIns, timeout: float | None = None, group_id: int | None = None) -> GetPropertiesRes:
        return GetPropertiesRes(Status(Code.OK), self.client.get_properties(ins.config))

    def get_parameters(self, ins: GetParametersIns, timeout: float | None = None, group_id: int | None = None) -> GetParametersRes:
        return GetParametersRes(Status(Code.OK), ndarrays_to_parameters(self.client.get_parameters(ins.config)))

    def fit(self, ins: FitIns, timeout: float | None = None, group_id: int | None = None) -> FitRes:
        arrays, num_examples, metrics = self.client.fit(parameters_to_ndarrays(ins.parameters), ins.config)
        return FitRes(Status(Code.OK), ndarrays_to_parameters(arrays), int(num_examples), metrics)

    def evaluate(self, ins: EvaluateIns, timeout: float | None = None, group_id: int | None = None) -> EvaluateRes:
        loss, num_examples, metrics = self.client.evaluate(parameters_to_ndarrays(ins.parameters), ins.config)
        return EvaluateRes(Status(Code.OK), float(loss), int(num_examples), metrics)

    def reconnect(self, ins: ReconnectIns, timeout: float | None = None, group_id: int | None = None) -> DisconnectRes:
        shutdown = getattr(self.client, "shutdown", None)
        if callable(shutdown):
            shutdown()
        return DisconnectRes(reason="shutdown")
