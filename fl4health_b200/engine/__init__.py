"""Execution engine: options, CUDA-graph step capture, fused flat optimizers, fast data staging."""

from fl4health_b200.engine.options import EngineOptions

__all__ = ["EngineOptions"]
