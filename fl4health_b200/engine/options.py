"""Per-client execution options (how the hooks are *run*, never what they compute)."""

from __future__ import annotations

import os
from dataclasses import dataclass

import torch


def _env_flag(name: str, default: bool) -> bool:
    raw = os.environ.get(name)
    if raw is None:
        return default
    return raw.strip().lower() not in ("0", "false", "off", "no", "")


@dataclass
class EngineOptions:
    """Defaults give exact-parity fp32 eager math with the flat arena + fused optimizer; the benchmark turns on
    bf16 autocast, channels-last and CUDA graphs.

    * ``arena``            re-home model state into a flat ``ParameterArena`` (fused exchange / optimizer)
    * ``fused_optimizer``  translate stock ``torch.optim.SGD/Adam/AdamW`` into one flat kernel launch per step
    * ``cuda_graphs``      capture ``train_step`` (+ loss/metric accumulation) into a CUDA graph after warm-up
    * ``amp_dtype``        autocast dtype for forward + loss (``None`` = fp32)
    * ``channels_last``    store 4-D parameters NHWC inside the arena and feed NHWC activations
    * ``master_weights``   with ``amp_dtype``: conv/linear parameters become ``amp_dtype`` views of an arena shadow
                           region (fp32 masters stay in the arena and are what is exchanged); the optimizer becomes one
                           multi-tensor launch that reads per-tensor low-precision gradients and writes master+shadow
    * ``table_grads``      (any precision) drop the flat gradient region: autograd assigns per-tensor gradients instead
                           of accumulating into arena views (one elementwise kernel per parameter and step saved), the
                           fused optimizer consumes them through a pointer table in one launch
    """

    arena: bool = True
    fused_optimizer: bool = True
    cuda_graphs: bool = False
    amp_dtype: torch.dtype | None = None
    channels_last: bool = False
    master_weights: bool = False
    table_grads: bool = False  # fp32 too: no flat gradient region, optimizer reads per-tensor grads (pointer table)
    graph_warmup_steps: int = 3
    step_reports: bool | None = None  # None = only when a reporter asks for per-step data

    @classmethod
    def from_env(cls) -> EngineOptions:
        amp = os.environ.get("FL4H_AMP", "").lower()
        amp_dtype = {"bf16": torch.bfloat16, "bfloat16": torch.bfloat16, "fp16": torch.float16}.get(amp)
        return cls(
            arena=_env_flag("FL4H_ARENA", True),
            fused_optimizer=_env_flag("FL4H_FUSED_OPT", True),
            cuda_graphs=_env_flag("FL4H_CUDA_GRAPHS", False),
            amp_dtype=amp_dtype,
            channels_last=_env_flag("FL4H_CHANNELS_LAST", False),
            master_weights=_env_flag("FL4H_MASTER_WEIGHTS", False),
            table_grads=_env_flag("FL4H_TABLE_GRADS", False),
        )
