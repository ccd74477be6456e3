from fl4health_b200.metrics.metrics import (
    F1,
    Accuracy,
    BalancedAccuracy,
    BinarySoftDiceCoefficient,
    RocAuc,
    SimpleMetric,
    TorchMetric,
)

from fl4health_b200.metrics.compound_metrics import EmaMetric, TransformsMetric
from fl4health_b200.metrics.efficient_metrics import BinaryDice, MultiClassDice

__all__ = ["F1", "Accuracy", "BalancedAccuracy", "BinarySoftDiceCoefficient", "RocAuc", "SimpleMetric", "TorchMetric", "EmaMetric", "TransformsMetric",
           "BinaryDice", "MultiClassDice"]
