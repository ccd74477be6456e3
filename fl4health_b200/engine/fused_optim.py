"""Flat-arena optimizers: one kernel launch per step over ``arena.flat[:trainable]``.

``FlatSGD`` / ``FlatAdamW`` are ``torch.optim.Optimizer`` subclasses (param groups, ``state_dict``, LR schedulers all
work) whose ``step`` is a single launch of ``ops.flat.sgd_step`` / ``adamw_step`` per contiguous parameter range
instead of the per-tensor ``foreach`` loops of the stock optimizers (SURVEY hot-op L6).  Two FL-specific terms are
folded into the same pass:

* ``anchor`` + ``mu``: the analytic gradient ``mu (w - w_t)`` of the FedProx/Ditto/MR-MTL drift penalty
  (``fl4health/losses/weight_drift_loss.py:57-64``) — no autograd graph over 3·L extra nodes;
* ``cv``: the SCAFFOLD correction ``c - c_i`` (``fl4health/clients/scaffold_client.py:187-197``), resident on the
  device instead of re-uploaded from NumPy every step.

``translate_optimizer`` converts a user's stock ``torch.optim.SGD/Adam/AdamW`` (as returned by ``get_optimizer``) when
its parameters all live in one arena.
"""

from __future__ import annotations

from typing import Any

import torch
from torch import nn
from torch.optim import Optimizer

from fl4health_b200.engine import streams
from fl4health_b200.ops import flat as F
from fl4health_b200.ops import multi_tensor as MT
from fl4health_b200.parallel.arena import ALIGN, ParameterArena, _round_up


def _capturing() -> bool:
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


def _group_ranges(arena: ParameterArena, params: list[nn.Parameter]) -> list[tuple[int, int]]:
    by_id = {id(p): name for name, p in arena.module.named_parameters(remove_duplicate=False)}
    names = []
    for p in params:
        if id(p) not in by_id:
            raise ValueError("optimizer parameter does not belong to the arena's module")
        name = arena.aliases.get(by_id[id(p)], by_id[id(p)])
        if arena.by_name[name].kind != "trainable":
            continue
        names.append(name)
    spans = sorted((arena.by_name[n].offset, _round_up(arena.by_name[n].end, ALIGN)) for n in set(names))
    merged: list[tuple[int, int]] = []
    for start, end in spans:
        if merged and start == merged[-1][1]:
            merged[-1] = (merged[-1][0], end)
        else:
            merged.append((start, end))
    return merged


class _FlatOptimizer(Optimizer):
    """Common machinery: ranges per group, device hyper-parameter blocks, extra fused terms."""

    def __init__(self, arena: ParameterArena, params: Any, defaults: dict[str, Any]) -> None:
        self.arena = arena
        super().__init__(params, defaults)
        # table mode (``arena.enable_compute_shadow``): no flat gradient region, per-tensor (bf16/fp32) gradients are
        # consumed through a pointer table by ONE multi-tensor launch per group (ops/csrc/mt_optim.cu)
        self.table_mode = arena.grad is None
        self._table_params: list[list[tuple[nn.Parameter, Any]]] = []
        if self.table_mode:
            by_id = {id(p): name for name, p in arena.module.named_parameters(remove_duplicate=False)}
            for group in self.param_groups:
                seen: set[str] = set()
                pairs = []
                for p in group["params"]:
                    if id(p) not in by_id:
                        raise ValueError("optimizer parameter does not belong to the arena's module")
                    name = arena.aliases.get(by_id[id(p)], by_id[id(p)])
                    if arena.by_name[name].kind == "trainable" and name not in seen:
                        seen.add(name)
                        pairs.append((p, arena.by_name[name]))
                self._table_params.append(pairs)
        self._ranges: list[list[tuple[int, int]]] = [_group_ranges(arena, g["params"]) for g in self.param_groups]
        self._hp: list[torch.Tensor] = [F.make_hyper_params(arena.device) for _ in self.param_groups]
        self._hp_cache: list[tuple | None] = [None for _ in self.param_groups]
        self.anchor: torch.Tensor | None = None  # FedProx / Ditto / MR-MTL w_t (arena offsets)
        self.mu: float = 0.0
        self.cv: torch.Tensor | None = None  # SCAFFOLD c - c_i (arena offsets)
        self.shadow: torch.Tensor | None = arena.shadow  # bf16 compute copy refreshed in the same pass
        self._grad_checked = 0

    # -- FL-specific fused terms ------------------------------------------------------------------------
    def set_drift_anchor(self, anchor: torch.Tensor | None, mu: float) -> None:
        self.anchor, self.mu = anchor, float(mu)

    def set_control_variate_correction(self, cv: torch.Tensor | None) -> None:
        self.cv = cv

    # -- hyper-parameter plumbing -----------------------------------------------------------------------
    def _hp_values(self, group: dict[str, Any]) -> dict[int, float]:
        raise NotImplementedError

    def sync_hyperparams(self) -> None:
        """Push changed host-side hyper-parameters (LR schedules, adaptive mu) into the device blocks.  Must be
        called outside CUDA-graph capture; the captured kernels read the block at replay time."""
        for idx, group in enumerate(self.param_groups):
            values = self._hp_values(group)
            values[F.HP_MU] = self.mu
            key = tuple(sorted(values.items()))
            if key != self._hp_cache[idx]:
                hp = self._hp[idx]
                for slot, value in values.items():
                    hp[slot] = value
                self._hp_cache[idx] = key

    def zero_grad(self, set_to_none: bool = False) -> None:  # noqa: ARG002 - grads are arena views, never dropped
        if self.table_mode:  # autograd then *assigns* fresh gradients: no zero-fill, no accumulate kernels
            for pairs in self._table_params:
                for p, _ in pairs:
                    p.grad = None
            return
        grad = self.arena.grad
        assert grad is not None
        covered = sum(end - start for ranges in self._ranges for start, end in ranges)
        if covered >= self.arena.trainable_padded:
            grad.zero_()
            return
        for ranges in self._ranges:  # an optimizer over a sub-module only clears its own gradients
            for start, end in ranges:
                grad[start:end].zero_()

    def _ensure_grad_views(self) -> None:
        """If something (e.g. ``zero_grad(set_to_none=True)`` on another handle) detached the gradient views, copy
        the stray grads into the arena and re-attach.  Checked on the first steps only."""
        if self.table_mode or self._grad_checked >= 2 or _capturing():
            return
        self._grad_checked += 1
        grad = self.arena.grad
        assert grad is not None
        params = dict(self.arena.module.named_parameters(remove_duplicate=False))
        for entry in self.arena.entries:
            if entry.kind != "trainable":
                continue
            p = params[entry.name]
            view = self.arena._shaped(grad, entry)
            if p.grad is None:
                view.zero_()
                p.grad = view
            elif p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
                p.grad = view

    def _table_entries(self, idx: int) -> list[MT.TableEntry]:
        entries = []
        for p, entry in self._table_params[idx]:
            g = p.grad
            if g is None:
                continue
            if g.stride() != p.stride():  # physical order must match the arena's (e.g. channels-last weights)
                g = torch.empty_like(p, dtype=g.dtype).copy_(g)
            entries.append(MT.TableEntry(g, entry.offset, entry.numel))
        return entries

    def _table_regions(self) -> tuple[torch.Tensor | None, torch.Tensor | None]:
        anchor = self.anchor if (self.anchor is not None and self.mu != 0.0) else None
        return anchor, self.cv

    def _slices(self, start: int, end: int) -> dict[str, torch.Tensor | None]:
        arena = self.arena
        assert arena.grad is not None
        return {
            "w": arena.flat[start:end],
            "g": arena.grad[start:end],
            "anchor": self.anchor[start:end] if (self.anchor is not None and self.mu != 0.0) else None,
            "cv": self.cv[start:end] if self.cv is not None else None,
            "shadow": self.shadow[start:end] if self.shadow is not None else None,
        }


class FlatSGD(_FlatOptimizer):
    def __init__(
        self,
        arena: ParameterArena,
        params: Any = None,
        lr: float = 1e-3,
        momentum: float = 0.0,
        dampening: float = 0.0,
        weight_decay: float = 0.0,
        nesterov: bool = False,
    ) -> None:
        if params is None:
            params = [p for p in arena.module.parameters() if p.requires_grad]
        defaults = dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay, nesterov=nesterov)
        super().__init__(arena, params, defaults)
        self.momentum_buffer = arena.companion(f"sgd_momentum_{id(self)}", trainable_only=True)
        for hp in self._hp:
            hp[F.HP_FIRST] = 1.0

    def _hp_values(self, group: dict[str, Any]) -> dict[int, float]:
        return {
            F.HP_LR: float(group["lr"]), F.HP_MOM: float(group["momentum"]), F.HP_DAMP: float(group["dampening"]),
            F.HP_WD: float(group["weight_decay"]), F.HP_NESTEROV: 1.0 if group["nesterov"] else 0.0,
        }

    @torch.no_grad()
    def step(self, closure: Any = None) -> Any:  # type: ignore[override]
        loss = closure() if closure is not None else None
        if streams.pending_count():  # gradients produced on the side stream (normally joined at the end of backward)
            streams.join()
        if not _capturing():
            self.sync_hyperparams()
            self._ensure_grad_views()
        if self.table_mode:
            anchor, cv = self._table_regions()
            for idx in range(len(self.param_groups)):
                MT.mt_step(self._table_entries(idx), False, self.arena.flat, self.momentum_buffer, None, self._hp[idx],
                           anchor, cv, self.shadow)
            return loss
        for idx, ranges in enumerate(self._ranges):
            for r, (start, end) in enumerate(ranges):
                s = self._slices(start, end)
                F.sgd_step(s["w"], s["g"], self.momentum_buffer[start:end], self._hp[idx], s["anchor"], s["cv"], s["shadow"],
                           more_ranges=r + 1 < len(ranges))
        return loss

    def state_dict(self) -> dict[str, Any]:  # type: ignore[override]
        base = super().state_dict()
        base["state"] = {
            "flat_momentum": self.momentum_buffer.detach().cpu().clone(),
            "first": [float(hp[F.HP_FIRST].item()) for hp in self._hp],
        }
        return base

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:  # type: ignore[override]
        state = state_dict.get("state", {})
        if "flat_momentum" in state:
            self.momentum_buffer.copy_(state["flat_momentum"].to(self.momentum_buffer.device))
            for hp, first in zip(self._hp, state.get("first", [])):
                hp[F.HP_FIRST] = first
        for group, saved in zip(self.param_groups, state_dict.get("param_groups", [])):
            group.update({k: v for k, v in saved.items() if k != "params"})
        self._hp_cache = [None for _ in self.param_groups]


class FlatAdamW(_FlatOptimizer):
    """Adam / AdamW (``decoupled_weight_decay`` selects which) over the flat arena."""

    def __init__(
        self,
        arena: ParameterArena,
        params: Any = None,
        lr: float = 1e-3,
        betas: tuple[float, float] = (0.9, 0.999),
        eps: float = 1e-8,
        weight_decay: float = 1e-2,
        decoupled_weight_decay: bool = True,
    ) -> None:
        if params is None:
            params = [p for p in arena.module.parameters() if p.requires_grad]
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(arena, params, defaults)
        self.decoupled = decoupled_weight_decay
        self.exp_avg = arena.companion(f"adam_m_{id(self)}", trainable_only=True)
        self.exp_avg_sq = arena.companion(f"adam_v_{id(self)}", trainable_only=True)

    def _hp_values(self, group: dict[str, Any]) -> dict[int, float]:
        return {
            F.HP_LR: float(group["lr"]), F.HP_B1: float(group["betas"][0]), F.HP_B2: float(group["betas"][1]),
            F.HP_EPS: float(group["eps"]), F.HP_WD: float(group["weight_decay"]),
        }

    @torch.no_grad()
    def step(self, closure: Any = None) -> Any:  # type: ignore[override]
        loss = closure() if closure is not None else None
        if streams.pending_count():
            streams.join()
        if not _capturing():
            self.sync_hyperparams()
            self._ensure_grad_views()
        if self.table_mode:
            anchor, _ = self._table_regions()
            for idx in range(len(self.param_groups)):
                MT.mt_step(self._table_entries(idx), True, self.arena.flat, self.exp_avg, self.exp_avg_sq, self._hp[idx],
                           anchor, None, self.shadow, self.decoupled)
            return loss
        for idx, ranges in enumerate(self._ranges):
            for r, (start, end) in enumerate(ranges):
                s = self._slices(start, end)
                hp = self._hp[idx]
                if r > 0:  # the step counter lives in the block and is ticked once per launch: undo for extra ranges
                    hp[F.HP_STEP] -= 1.0
                F.adamw_step(s["w"], s["g"], self.exp_avg[start:end], self.exp_avg_sq[start:end], hp, s["anchor"],
                             s["shadow"], self.decoupled)
        return loss

    def state_dict(self) -> dict[str, Any]:  # type: ignore[override]
        base = super().state_dict()
        base["state"] = {
            "flat_exp_avg": self.exp_avg.detach().cpu().clone(),
            "flat_exp_avg_sq": self.exp_avg_sq.detach().cpu().clone(),
            "step": [float(hp[F.HP_STEP].item()) for hp in self._hp],
        }
        return base

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:  # type: ignore[override]
        state = state_dict.get("state", {})
        if "flat_exp_avg" in state:
            self.exp_avg.copy_(state["flat_exp_avg"].to(self.exp_avg.device))
            self.exp_avg_sq.copy_(state["flat_exp_avg_sq"].to(self.exp_avg_sq.device))
            for hp, step in zip(self._hp, state.get("step", [])):
                hp[F.HP_STEP] = step
        for group, saved in zip(self.param_groups, state_dict.get("param_groups", [])):
            group.update({k: v for k, v in saved.items() if k != "params"})
        self._hp_cache = [None for _ in self.param_groups]


def translate_optimizer(optimizer: Optimizer, arena: ParameterArena) -> Optimizer:
    """Stock SGD / Adam / AdamW -> flat fused equivalent.  Anything else (or exotic flags) is returned unchanged."""
    if isinstance(optimizer, _FlatOptimizer) or getattr(optimizer, "fl4h_keep_stock", False):
        return optimizer  # (ZeRO-1 shards keep the stock optimizer: the flat companions are arena-length by construction)
    kind = type(optimizer)
    trainable_arena = arena.grad is not None or arena.shadow is not None or getattr(arena, "table_gradients", False)
    if kind not in (torch.optim.SGD, torch.optim.Adam, torch.optim.AdamW) or not trainable_arena:
        return optimizer  # (an arena attached without gradients holds an evaluation-only / frozen model)
    module_params = {id(p) for p in arena.module.parameters()}
    groups = []
    for group in optimizer.param_groups:
        if any(id(p) not in module_params for p in group["params"]):
            return optimizer
        if group.get("maximize") or group.get("amsgrad") or group.get("differentiable"):
            return optimizer
        groups.append(group)
    if len(optimizer.state) > 0:  # already stepped: keep the user's state semantics
        return optimizer

    def make_groups(keys: list[str]) -> list[dict[str, Any]]:
        return [{"params": g["params"], **{k: g[k] for k in keys}} for g in groups]

    if kind is torch.optim.SGD:
        first = groups[0]
        return FlatSGD(
            arena, make_groups(["lr", "momentum", "dampening", "weight_decay", "nesterov"]), lr=first["lr"],
            momentum=first["momentum"], dampening=first["dampening"], weight_decay=first["weight_decay"],
            nesterov=first["nesterov"],
        )
    first = groups[0]
    decoupled = kind is torch.optim.AdamW or bool(first.get("decoupled_weight_decay", False))
    return FlatAdamW(
        arena, make_groups(["lr", "betas", "eps", "weight_decay"]), lr=first["lr"], betas=first["betas"],
        eps=first["eps"], weight_decay=first["weight_decay"], decoupled_weight_decay=decoupled,
    )
