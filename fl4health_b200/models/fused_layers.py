"""Drop-in layers backed by the fused sm_100a kernels (same parameters / buffers / state-dict keys as the stock
``torch.nn`` layers they subclass, so checkpoints and parameter exchange are unaffected)."""

from __future__ import annotations

import torch
from torch import nn

from fl4health_b200.ops.bn_act import batch_norm_act


class BatchNormAct2d(nn.BatchNorm2d):
    """``relu(bn(x) + residual)`` in two kernel launches forward / two backward on channels-last CUDA tensors
    (``ops/csrc/bn_act.cu``); exactly ``nn.BatchNorm2d`` (+ add + relu) everywhere else."""

    def __init__(self, num_features: int, eps: float = 1e-5, momentum: float | None = 0.1, affine: bool = True,
                 track_running_stats: bool = True, relu: bool = True, device=None, dtype=None) -> None:  # noqa: ANN001
        super().__init__(num_features, eps, momentum, affine, track_running_stats, device=device, dtype=dtype)
        self.relu = relu

    def forward(self, x: torch.Tensor, residual: torch.Tensor | None = None) -> torch.Tensor:  # type: ignore[override]
        self._check_input_dim(x)
        return batch_norm_act(
            x, self.weight, self.bias, self.running_mean if self.track_running_stats else None,
            self.running_var if self.track_running_stats else None,
            self.num_batches_tracked if self.track_running_stats else None, self.training, self.momentum, self.eps,
            residual=residual, relu=self.relu,
        )

    def extra_repr(self) -> str:
        return super().extra_repr() + f", relu={self.relu}"


def bn_act(bn: nn.Module, x: torch.Tensor, residual: torch.Tensor | None = None, relu: bool = True) -> torch.Tensor:
    """Apply ``bn`` then (+residual)(relu).  Works for any normalisation module (e.g. the GroupNorm a DP validator
    swapped in); the fused path is taken when ``bn`` is a ``BatchNormAct2d``."""
    if isinstance(bn, BatchNormAct2d):
        assert bn.relu == relu
        return bn(x, residual)
    out = bn(x)
    if residual is not None:
        out = out + residual
    return torch.relu(out) if relu else out
