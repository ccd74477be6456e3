"""``NumPyClient`` and the adapter that speaks the ``*Ins`` / ``*Res`` vocabulary on its behalf."""

from __future__ import annotations

from ..common.parameter import ndarrays_to_parameters, parameters_to_ndarrays
from ..common.typing import (
    Code,
    Config,
    EvaluateIns,
    EvaluateRes,
    FitIns,
    FitRes,
    GetParametersIns,
    GetParametersRes,
    GetPropertiesIns,
    GetPropertiesRes,
    NDArrays,
    Scalar,
    Status,
)


class NumPyClient:
    """User-facing base class: every method works on lists of NumPy arrays."""

    def get_properties(self, config: Config) -> dict[str, Scalar]:
        return {}

    def get_parameters(self, config: dict[str, Scalar]) -> NDArrays:
        return []

    def fit(self, parameters: NDArrays, config: dict[str, Scalar]) -> tuple[NDArrays, int, dict[str, Scalar]]:
        return [], 0, {}

    def evaluate(self, parameters: NDArrays, config: dict[str, Scalar]) -> tuple[float, int, dict[str, Scalar]]:
        return 0.0, 0, {}

    def to_client(self) -> "Client":
        return Client(self)


_OK = Status(code=Code.OK, message="Success")


class Client:
    """Message-level adapter around a :class:`NumPyClient` (deserialise -> call -> serialise)."""

    def __init__(self, numpy_client: NumPyClient) -> None:
        self.numpy_client = numpy_client

    def get_properties(self, ins: GetPropertiesIns) -> GetPropertiesRes:
        return GetPropertiesRes(status=_OK, properties=self.numpy_client.get_properties(config=ins.config))

    def get_parameters(self, ins: GetParametersIns) -> GetParametersRes:
        arrays = self.numpy_client.get_parameters(config=ins.config)
        return GetParametersRes(status=_OK, parameters=ndarrays_to_parameters(arrays))

    def fit(self, ins: FitIns) -> FitRes:
        arrays, num_examples, metrics = self.numpy_client.fit(parameters_to_ndarrays(ins.parameters), ins.config)
        return FitRes(status=_OK, parameters=ndarrays_to_parameters(arrays), num_examples=num_examples, metrics=metrics)

    def evaluate(self, ins: EvaluateIns) -> EvaluateRes:
        loss, num_examples, metrics = self.numpy_client.evaluate(parameters_to_ndarrays(ins.parameters), ins.config)
        return EvaluateRes(status=_OK, loss=float(loss), num_examples=num_examples, metrics=metrics)
