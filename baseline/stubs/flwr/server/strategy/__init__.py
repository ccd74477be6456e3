from .fedavg import FedAvg  # noqa: F401
from .strategy import Strategy  # noqa: F401
