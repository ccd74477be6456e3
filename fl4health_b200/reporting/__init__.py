from fl4health_b200.reporting.json_reporter import JsonReporter
from fl4health_b200.reporting.wandb_reporter import WandBReporter, WandBStepType

__all__ = ["JsonReporter", "WandBReporter", "WandBStepType"]
