from fl4health_b200.metrics.metrics import (
    F1,
    Accuracy,
    BalancedAccuracy,
    BinarySoftDiceCoefficient,
    RocAuc,
    SimpleMetric,
    TorchMetric,
)

__all__ = ["F1", "Accuracy", "BalancedAccuracy", "BinarySoftDiceCoefficient", "RocAuc", "SimpleMetric", "TorchMetric"]
