"""Weight-drift (proximal) penalty ``(mu/2) * sum_l ||w_l - w_l^ref||^2``.

Parity: ``fl4health/losses/weight_drift_loss.py:5-64`` (same call signature).  The reference builds one
``linalg.norm`` + ``pow`` per layer and a ``stack().sum()`` — ~3 autograd nodes per layer, every step.
When the model lives in a ``ParameterArena`` and the reference tensors are views of one arena-shaped region, the value
is ONE flat reduction kernel and the gradient ``mu (w - w_ref)`` ONE fused elementwise pass whose per-parameter views
are handed to autograd (no per-layer kernels).  ``BasicClient``-derived clients with a fused flat optimizer go one
step further and fold the gradient into the optimizer kernel, so this loss only reports the value.
"""

from __future__ import annotations

from typing import Any

import torch
from torch import nn

from fl4health_b200.ops import flat as flat_ops
from fl4health_b200.parallel.arena import ParameterArena, arena_of


def flat_reference_region(arena: ParameterArena, constraint_tensors: list[torch.Tensor]) -> torch.Tensor | None:
    """If ``constraint_tensors`` are exactly the per-parameter views of one arena-shaped region, return that region."""
    if not constraint_tensors:
        return None
    params = [(name, p) for name, p in arena.module.named_parameters()]
    if len(params) != len(constraint_tensors):
        return None
    first_entry = arena.by_name[arena.aliases.get(params[0][0], params[0][0])]
    base_ptr = constraint_tensors[0].data_ptr() - first_entry.offset * 4
    for (name, _), ref in zip(params, constraint_tensors):
        entry = arena.by_name[arena.aliases.get(name, name)]
        if ref.dtype != torch.float32 or ref.data_ptr() != base_ptr + entry.offset * 4:
            return None
    for region in arena.regions.values():
        if region.data_ptr() == base_ptr:
            return region
    return None


class _FlatDrift(torch.autograd.Function):
    """value = mu/2 * ||w - a||^2 over the trainable prefix of the arena; grads are views of ONE fused delta buffer."""

    @staticmethod
    def forward(ctx: Any, arena: ParameterArena, anchor: torch.Tensor, weight: float, *params: torch.Tensor) -> torch.Tensor:  # type: ignore[override]
        n = arena.trainable_padded
        ctx.arena, ctx.anchor, ctx.weight = arena, anchor, weight
        value = flat_ops.sq_diff_sum(arena.flat[:n], anchor[:n])
        return (value * (weight / 2.0)).reshape(())

    @staticmethod
    def backward(ctx: Any, grad_output: torch.Tensor) -> tuple:  # type: ignore[override]
        arena: ParameterArena = ctx.arena
        n = arena.trainable_padded
        delta = (arena.flat[:n] - ctx.anchor[:n]) * (ctx.weight * grad_output)
        grads: list[torch.Tensor | None] = []
        for name, p in arena.module.named_parameters():
            entry = arena.by_name[arena.aliases.get(name, name)]
            grads.append(arena._shaped(delta, entry) if (p.requires_grad and entry.kind == "trainable") else None)
        return (None, None, None, *grads)


class WeightDriftLoss(nn.Module):
    def __init__(self, device: torch.device) -> None:
        super().__init__()
        self.device = device

    def _compute_weight_difference_inner_product(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        return torch.pow(torch.linalg.norm(x - y), 2.0)

    def forward(self, target_model: nn.Module, constraint_tensors: list[torch.Tensor], weight: float) -> torch.Tensor:
        model_weights = list(target_model.parameters())
        assert len(constraint_tensors) == len(model_weights)
        assert len(model_weights) > 0
        arena = arena_of(target_model)
        if arena is not None and arena.trainable_numel > 0:
            region = flat_reference_region(arena, constraint_tensors)
            if region is not None:
                return _FlatDrift.apply(arena, region, float(weight), *model_weights)
        constraint_tensors = [t.to(self.device) for t in constraint_tensors]
        inner_products = [
            self._compute_weight_difference_inner_product(ref, w) for ref, w in zip(constraint_tensors, model_weights)
        ]
        # the 1/2 makes the gradient exactly weight * (w - ref)
        return (weight / 2.0) * torch.stack(inner_products).sum()
