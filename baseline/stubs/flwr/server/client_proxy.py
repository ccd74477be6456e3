"""Abstract server-side handle of a client."""

from __future__ import annotations

from abc import ABC, abstractmethod

from ..common.typing import (
    DisconnectRes,
    EvaluateIns,
    EvaluateRes,
    FitIns,
    FitRes,
    GetParametersIns,
    GetParametersRes,
    GetPropertiesIns,
    GetPropertiesRes,
    Properties,
    ReconnectIns,
)


class ClientProxy(ABC):
    node_id: int

    def __init__(self, cid: str) -> None:
        self.cid = cid
        self.properties: Properties = {}

    @abstractmethod
    def get_properties(self, ins: GetPropertiesIns, timeout: float | None, group_id: int | None) -> GetPropertiesRes: ...

    @abstractmethod
    def get_parameters(self, ins: GetParametersIns, timeout: float | None, group_id: int | None) -> GetParametersRes: ...

    @abstractmethod
    def fit(self, ins: FitIns, timeout: float | None, group_id: int | None) -> FitRes: ...

    @abstractmethod
    def evaluate(self, ins: EvaluateIns, timeout: float | None, group_id: int | None) -> EvaluateRes: ...

    @abstractmethod
    def reconnect(self, ins: ReconnectIns, timeout: float | None, group_id: int | None) -> DisconnectRes: ...
