from .logger import configure, log  # noqa: F401
from .parameter import bytes_to_ndarray, ndarray_to_bytes, ndarrays_to_parameters, parameters_to_ndarrays  # noqa: F401
from .typing import (  # noqa: F401
    Code,
    Config,
    DisconnectRes,
    EvaluateIns,
    EvaluateRes,
    FitIns,
    FitRes,
    GetParametersIns,
    GetParametersRes,
    GetPropertiesIns,
    GetPropertiesRes,
    Metrics,
    MetricsAggregationFn,
    NDArray,
    NDArrays,
    Parameters,
    Properties,
    ReconnectIns,
    Scalar,
    Status,
)

GRPC_MAX_MESSAGE_LENGTH = 536_870_912
