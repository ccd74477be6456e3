"""Flat parameter arena: the memory layout everything else in the engine is built around.

A model's exchangeable floating-point state (parameters *and* float buffers such as BatchNorm running statistics) is
re-homed into ONE contiguous fp32 buffer per rank; ``nn.Parameter.data`` / buffers become views into it.  Consequences:

* parameter exchange is a single contiguous buffer operation (one fused kernel / one collective) instead of the
  reference's per-layer ``val.cpu().numpy()`` / ``torch.tensor(v)`` loops (``full_exchanger.py:30,45-47``);
* the local optimizer is one multi-tensor kernel over ``flat[:trainable]`` + ``grad[:trainable]`` with the FedProx
  drift term and SCAFFOLD correction folded in;
* companion regions (gradient, momentum, FedProx anchor ``w_t``, SCAFFOLD ``c``/``c_i``/``c-c_i``, server moments,
  bf16 compute shadow) share the same offsets, so "slices" (FedPer base module, FedBN exclusions, ...) are just
  offset ranges;
* on multi-GPU runs the arena is allocated from peer-mapped symmetric memory so other ranks' kernels can read/write it
  over NVLink directly.

Layout: ``[trainable params | frozen params | float buffers]``, every entry aligned to ``ALIGN`` elements; the list
order handed to strategies still follows ``state_dict()`` order (views may point anywhere in the flat buffer).
Integer buffers (``num_batches_tracked``) are not in the arena; they travel as separate tiny tensors.
"""

from __future__ import annotations

from collections import OrderedDict
from collections.abc import Callable, Iterable
from dataclasses import dataclass

import torch
from torch import nn

from fl4health_b200.common.typing import NDArrays

import weakref

_ARENAS: "weakref.WeakKeyDictionary[nn.Module, ParameterArena]" = weakref.WeakKeyDictionary()

ALIGN = 32  # elements (128 B of fp32): keeps every entry 16B-vectorizable in fp32 *and* bf16


@dataclass(frozen=True)
class ArenaEntry:
    name: str
    offset: int
    numel: int
    shape: tuple[int, ...]
    kind: str  # "trainable" | "frozen" | "buffer"
    dtype: torch.dtype  # original dtype of the tensor (restored on export)
    nhwc: bool = False  # 4-D tensor stored channels-last inside the flat buffer

    @property
    def end(self) -> int:
        return self.offset + self.numel


def _round_up(n: int, multiple: int) -> int:
    return (n + multiple - 1) // multiple * multiple


Allocator = Callable[[int, torch.dtype, torch.device], torch.Tensor]


def _default_allocator(numel: int, dtype: torch.dtype, device: torch.device) -> torch.Tensor:
    return torch.zeros(numel, dtype=dtype, device=device)


class ParameterArena:
    """Flat fp32 home for a module's float state, plus named companion regions with identical offsets."""

    def __init__(
        self,
        module: nn.Module,
        device: torch.device | str | None = None,
        allocator: Allocator | None = None,
        with_grad: bool = True,
        channels_last: bool = False,
    ) -> None:
        self.module = module
        self.channels_last = channels_last
        first = next(iter(module.state_dict().values()), None)
        self.device = torch.device(device) if device is not None else (first.device if first is not None else torch.device("cpu"))
        self._alloc: Allocator = allocator or _default_allocator
        self.entries: list[ArenaEntry] = []
        self.by_name: dict[str, ArenaEntry] = {}
        self.int_state: OrderedDict[str, torch.Tensor] = OrderedDict()
        self.state_keys: list[str] = []
        self.regions: dict[str, torch.Tensor] = {}
        self._build_layout()
        self.flat = self._alloc(self.total, torch.float32, self.device)
        self.int_flat: torch.Tensor | None = None
        self.anchor_on_pull = False  # drift-constrained clients: every full pull also writes regions["drift_anchor"]
        self.anchor_fresh = False    # set by the fused pull, consumed by snapshot_drift_anchor
        self._views_cache: dict[tuple[int, int], NDArrays] = {}
        self._rehome()
        self.grad: torch.Tensor | None = None
        self._subset_views_cache: dict[tuple, NDArrays] = {}
        self.table_gradients = False  # True: trainable, but gradients are per-tensor (no flat region); see use_table_gradients
        self.shadow: torch.Tensor | None = None  # bf16 compute copy (see enable_compute_shadow)
        self.shadow_names: set[str] = set()
        if with_grad and self.trainable_numel > 0:
            self.grad = self._alloc(self.trainable_padded, torch.float32, self.device)
            self._attach_grads()
        _ARENAS[module] = self

    # ------------------------------------------------------------------------------------------------------
    def _build_layout(self) -> None:
        params = dict(self.module.named_parameters(remove_duplicate=False))
        state = self.module.state_dict(keep_vars=True)
        self.state_keys = list(state.keys())
        groups: dict[str, list[tuple[str, torch.Tensor]]] = {"trainable": [], "frozen": [], "buffer": []}
        seen: dict[int, str] = {}
        self.aliases: dict[str, str] = {}
        for name, tensor in state.items():
            if id(tensor) in seen:  # tied weights: one home, several names
                self.aliases[name] = seen[id(tensor)]
                continue
            seen[id(tensor)] = name
            if not tensor.is_floating_point():
                self.int_state[name] = tensor
                continue
            if name in params:
                groups["trainable" if params[name].requires_grad else "frozen"].append((name, tensor))
            else:
                groups["buffer"].append((name, tensor))
        offset = 0
        for kind in ("trainable", "frozen", "buffer"):
            for name, tensor in groups[kind]:
                nhwc = self.channels_last and tensor.dim() == 4
                entry = ArenaEntry(name, offset, tensor.numel(), tuple(tensor.shape), kind, tensor.dtype, nhwc)
                self.entries.append(entry)
                self.by_name[name] = entry
                offset = _round_up(offset + tensor.numel(), ALIGN)
            if kind == "trainable":
                self.trainable_padded = offset
        self.total = max(offset, ALIGN)
        self.trainable_numel = sum(e.numel for e in self.entries if e.kind == "trainable")
        self.numel = sum(e.numel for e in self.entries)

    @staticmethod
    def _shaped(base: torch.Tensor, entry: ArenaEntry) -> torch.Tensor:
        chunk = base[entry.offset : entry.end]
        if entry.nhwc:
            n, c, h, w = entry.shape
            return chunk.view(n, h, w, c).permute(0, 3, 1, 2)  # logical NCHW, physical NHWC
        return chunk.view(entry.shape)

    def _find_owner(self, dotted: str) -> tuple[nn.Module, str]:
        owner: nn.Module = self.module
        parts = dotted.split(".")
        for part in parts[:-1]:
            owner = getattr(owner, part)
        return owner, parts[-1]

    def _rehome(self) -> None:
        state = self.module.state_dict(keep_vars=True)
        with torch.no_grad():
            for entry in self.entries:
                src = state[entry.name]
                view = self._shaped(self.flat, entry)
                view.copy_(src.detach().to(device=self.device, dtype=torch.float32))
                if isinstance(src, nn.Parameter):
                    src.data = view
                else:
                    owner, leaf = self._find_owner(entry.name)
                    owner._buffers[leaf] = _own_version_alias(view)
            # integer buffers (num_batches_tracked ...) live in ONE int64 buffer so they aggregate with one op
            int_sizes = [(name, t.numel()) for name, t in self.int_state.items() if t.dtype == torch.int64]
            if int_sizes:
                # same allocator as the float arena: in SPMD mode the counters sit in symmetric memory, so the fused
                # aggregate kernel reduces them through peer loads in the same launch (no separate collective)
                self.int_flat = self._alloc(sum(n for _, n in int_sizes), torch.int64, self.device)
                self.int_flat.zero_()
                cursor = 0
                for name, numel in int_sizes:
                    old = self.int_state[name]
                    view = self.int_flat[cursor : cursor + numel].view(old.shape)
                    view.copy_(old.to(self.device))
                    owner, leaf = self._find_owner(name)
                    if leaf in owner._buffers:
                        owner._buffers[leaf] = view
                    self.int_state[name] = view
                    cursor += numel
            for name, tensor in list(self.int_state.items()):
                if tensor.device != self.device:
                    owner, leaf = self._find_owner(name)
                    moved = tensor.to(self.device)
                    if leaf in owner._buffers:
                        owner._buffers[leaf] = moved
                    self.int_state[name] = moved

    def _attach_grads(self) -> None:
        assert self.grad is not None
        params = dict(self.module.named_parameters(remove_duplicate=False))
        for entry in self.entries:
            if entry.kind == "trainable":
                params[entry.name].grad = self._shaped(self.grad, entry)

    # ------------------------------------------------------------------------------------------------------
    def enable_compute_shadow(
        self, dtype: torch.dtype = torch.bfloat16, module_types: tuple[type, ...] | None = None
    ) -> torch.Tensor:
        """Master-weight mode: the parameters of GEMM-shaped layers (conv / linear by default) become ``dtype`` views
        of a *shadow* region with the arena's offsets; ``flat`` keeps the fp32 masters that are exchanged, aggregated
        and checkpointed.  The forward then needs no per-step weight casts, autograd produces ``dtype`` gradients
        (no cast-back kernels), and gradients are no longer accumulated into a flat fp32 region: ``.grad`` is None
        before each backward, so autograd assigns instead of adding, and the multi-tensor optimizer kernel
        (``ops/csrc/mt_optim.cu``) consumes the per-tensor gradients through a pointer table while writing both the
        master and the shadow.  Normalisation parameters stay fp32 views of the master.

        Whoever writes ``flat`` outside the optimizer (parameter pull, aggregation broadcast) must call
        ``refresh_shadow``; ``load_ndarrays`` does so itself."""
        if self.shadow is not None:
            return self.shadow
        if module_types is None:
            module_types = (nn.modules.conv._ConvNd, nn.Linear)
        self.shadow = self._alloc(self.total, dtype, self.device)
        params = dict(self.module.named_parameters(remove_duplicate=False))
        selected: set[int] = set()
        for mod in self.module.modules():
            if isinstance(mod, module_types):
                selected.update(id(p) for p in mod.parameters(recurse=False))
        for entry in self.entries:
            if entry.kind == "buffer":
                continue
            param = params[entry.name]
            if id(param) in selected:
                param.data = self._shaped(self.shadow, entry)
                self.shadow_names.add(entry.name)
            param.grad = None
        self.grad = None  # table mode: gradients live wherever autograd allocates them
        self.refresh_shadow()
        return self.shadow

    def use_table_gradients(self) -> None:
        """Drop the flat gradient region: ``.grad`` is None before every backward, so autograd ASSIGNS each parameter's
        gradient (no zero-fill and no accumulate kernel per parameter) and the multi-tensor optimizer reads the
        per-tensor gradients through a pointer table (``ops/csrc/mt_optim.cu``).  What ``enable_compute_shadow`` does
        for bf16 master weights, available to plain fp32 training too (``EngineOptions.table_grads``)."""
        for param in self.module.parameters():
            param.grad = None
        self.grad = None
        self.table_gradients = True

    @property
    def params_end(self) -> int:
        ends = [_round_up(e.end, ALIGN) for e in self.entries if e.kind != "buffer"]
        return max(ends) if ends else 0

    def refresh_shadow(self) -> None:
        if self.shadow is None:
            return
        from fl4health_b200.ops import flat as flat_ops

        end = self.params_end
        with torch.no_grad():
            flat_ops.cast_bf16(self.flat[:end], self.shadow[:end]) if self.shadow.dtype == torch.bfloat16 else self.shadow[
                :end
            ].copy_(self.flat[:end])

    def view(self, name: str, region: torch.Tensor | None = None) -> torch.Tensor:
        entry = self.by_name[self.aliases.get(name, name)]
        base = self.flat if region is None else region
        return self._shaped(base, entry)

    def companion(self, name: str, trainable_only: bool = False, dtype: torch.dtype = torch.float32) -> torch.Tensor:
        """Get-or-create a zero-initialised region with the arena's offsets (anchor, momentum, variates, ...)."""
        if name not in self.regions:
            numel = self.trainable_padded if trainable_only else self.total
            self.regions[name] = self._alloc(max(numel, ALIGN), dtype, self.device)
        return self.regions[name]

    def zero_grad(self) -> None:
        if self.grad is not None:
            self.grad.zero_()

    def trainable(self, region: torch.Tensor | None = None) -> torch.Tensor:
        base = self.flat if region is None else region
        return base[: self.trainable_padded]

    # ------------------------------------------------------------------------------------------------------
    def ndarrays(self, names: Iterable[str] | None = None, region: torch.Tensor | None = None) -> NDArrays:
        """state_dict-ordered list of views (``FullParameterExchanger`` order, ``full_exchanger.py:30``)."""
        base = self.flat if region is None else region
        if names is None:
            cache_key = (base.data_ptr(), base.numel())
            cached = self._views_cache.get(cache_key)
            # same address + size + dtype while the cache keeps the old storage alive => same memory: reuse the views
            if cached is not None and cached.flat.dtype == base.dtype and cached.flat.device == base.device:
                fresh = NDArrays(cached, flat=base, layout=self)  # shallow copy: callers may mutate the list
                fresh.int_flat = self.int_flat if region is None else None
                return fresh
        keys = list(names) if names is not None else self.state_keys
        if names is not None:
            # a fixed subset (FedPer / FedRep / FedBN exchange the same names every round): the views are built once
            subset_key = (base.data_ptr(), base.numel(), tuple(keys))
            cached = self._subset_views_cache.get(subset_key)
            if cached is not None and cached.subset_flat.dtype == base.dtype and cached.subset_flat.device == base.device:
                fresh = NDArrays(cached)
                fresh.subset_flat, fresh.subset_layout, fresh.subset_names = base, self, cached.subset_names
                return fresh
        out = NDArrays()
        for key in keys:
            key = self.aliases.get(key, key)
            if key in self.by_name:
                out.append(self.view(key, region))
            else:
                out.append(self.int_state[key])
        if names is None:
            out.flat = base
            out.layout = self
            out.int_flat = self.int_flat if region is None else None
            if len(self._views_cache) > 8:
                self._views_cache.pop(next(iter(self._views_cache)))
            self._views_cache[(base.data_ptr(), base.numel())] = NDArrays(out, flat=base, layout=self)
        else:  # a named subset of the arena: tagged so that it can ride the whole-arena collectives / copies
            out.subset_flat, out.subset_layout, out.subset_names = base, self, tuple(keys)
            if len(self._subset_views_cache) > 8:
                self._subset_views_cache.pop(next(iter(self._subset_views_cache)))
            keep = NDArrays(out)
            keep.subset_flat, keep.subset_layout, keep.subset_names = base, self, tuple(keys)
            self._subset_views_cache[(base.data_ptr(), base.numel(), tuple(keys))] = keep
        return out

    def subset_plan(self, names: tuple[str, ...]) -> tuple[list[tuple[int, int]], list[int]]:
        """For a named subset: the arena element ranges its float entries cover (adjacent entries merged -- a FedPer
        feature extractor is two ranges: its parameters and its BatchNorm buffers) and the list positions of its
        integer entries.  Cached per subset."""
        cache = self.__dict__.setdefault("_subset_plans", {})
        plan = cache.get(names)
        if plan is None:
            spans, int_positions = [], []
            for position, name in enumerate(names):
                key = self.aliases.get(name, name)
                if key in self.by_name:
                    entry = self.by_name[key]
                    spans.append((entry.offset, _round_up(entry.end, ALIGN)))
                else:
                    int_positions.append(position)
            merged: list[tuple[int, int]] = []
            for start, end in sorted(set(spans)):
                if merged and start <= merged[-1][1]:
                    merged[-1] = (merged[-1][0], max(merged[-1][1], end))
                else:
                    merged.append((start, end))
            plan = cache[names] = ([(a, min(b, self.flat.numel())) for a, b in merged], int_positions)
        return plan

    def load_ndarrays(self, arrays: list, names: Iterable[str] | None = None) -> None:
        """Copy a list of arrays into the arena.  A single flat copy when the source is arena-shaped."""
        keys = list(names) if names is not None else self.state_keys
        src_flat = getattr(arrays, "flat", None)
        src_layout = getattr(arrays, "layout", None)
        self.anchor_fresh = False  # only a fused full pull (below) may vouch for the anchor region
        with torch.no_grad():
            if names is None and src_flat is not None and isinstance(src_layout, ParameterArena) and src_layout.same_layout(self):
                fused_pull = self._fused_pull(src_flat)
                if not fused_pull and src_flat.data_ptr() != self.flat.data_ptr():
                    self.flat.copy_(src_flat, non_blocking=True)
                src_int = getattr(arrays, "int_flat", None)
                all_flat = self.int_flat is not None and all(t.dtype == torch.int64 for t in self.int_state.values())
                if all_flat and src_int is not None and src_int.numel() == self.int_flat.numel():
                    if src_int.data_ptr() != self.int_flat.data_ptr():
                        self.int_flat.copy_(src_int, non_blocking=True)
                else:
                    for idx in self._int_positions():
                        dst_int = self.int_state[self.aliases.get(keys[idx], keys[idx])]
                        dst_int.copy_(_as_tensor(arrays[idx], self.device).to(dst_int.dtype).reshape(dst_int.shape))
                if not fused_pull:
                    self.refresh_shadow()
                return
            assert len(keys) == len(arrays), f"expected {len(keys)} arrays, received {len(arrays)}"
            subset_flat = getattr(arrays, "subset_flat", None)
            subset_layout = getattr(arrays, "subset_layout", None)
            if (subset_flat is not None and isinstance(subset_layout, ParameterArena) and subset_layout.same_layout(self)
                    and getattr(arrays, "subset_names", None) == tuple(keys) and subset_flat.numel() == self.flat.numel()
                    and subset_flat.dtype == self.flat.dtype and subset_flat.device == self.flat.device):
                # the payload is an arena-shaped buffer + the names that matter in it: copy the few ranges they cover
                ranges, int_positions = self.subset_plan(tuple(keys))
                for start, end in ranges:
                    self.flat[start:end].copy_(subset_flat[start:end], non_blocking=True)
                if int_positions:
                    targets = [self.int_state[self.aliases.get(keys[i], keys[i])] for i in int_positions]
                    torch._foreach_copy_(targets, [_as_tensor(arrays[i], self.device).to(t.dtype).reshape(t.shape)
                                                   for i, t in zip(int_positions, targets)], non_blocking=True)
                self.refresh_shadow()
                return
            destinations, sources = [], []
            for key, arr in zip(keys, arrays):
                key = self.aliases.get(key, key)
                t = _as_tensor(arr, self.device)
                if key in self.by_name:
                    dst = self.view(key)
                    assert dst.shape == t.shape, f"shape mismatch for {key}: {tuple(dst.shape)} vs {tuple(t.shape)}"
                    destinations.append(dst)
                    sources.append(t if t.dtype == dst.dtype else t.to(dst.dtype))
                else:
                    dst_int = self.int_state[key]
                    dst_int.copy_(t.to(dst_int.dtype).reshape(dst_int.shape))
            if destinations:  # partial payloads (FedPer, FedRep, layer exchange): a couple of batched launches, not one per tensor
                torch._foreach_copy_(destinations, sources, non_blocking=True)
            self.refresh_shadow()

    def _fused_pull(self, src_flat: torch.Tensor) -> bool:
        """The server->client pull as ONE kernel (SURVEY C1 receiver side): a single read of the landed global buffer
        writes the fp32 masters, the bf16 compute shadow and — when a drift-constrained client registered
        ``anchor_on_pull`` — the FedProx anchor ``w_t`` (replaces copy + cast + anchor-snapshot launches; reference:
        fl4health/parameter_exchange/full_exchanger.py:45-47 + fl4health/clients/fed_prox_client.py:18-20)."""
        if not self.flat.is_cuda or src_flat.dtype != torch.float32 or src_flat.numel() != self.flat.numel():
            return False
        if src_flat.data_ptr() == self.flat.data_ptr() or self.flat.numel() % 4:
            return False
        shadow = self.shadow if (self.shadow is not None and self.shadow.dtype == torch.bfloat16
                                 and self.shadow.numel() >= self.flat.numel()) else None
        anchor = self.regions.get("drift_anchor") if self.anchor_on_pull else None
        if anchor is not None and anchor.numel() < self.flat.numel():
            anchor = None
        if shadow is None and anchor is None:
            return False  # a plain copy is already one launch
        from fl4health_b200.ops import flat as flat_ops

        flat_ops.bcast_unpack(src_flat, w=self.flat, anchor=anchor, shadow=shadow)
        self.anchor_fresh = anchor is not None
        if self.shadow is not None and shadow is None:
            self.refresh_shadow()
        return True

    def _int_positions(self) -> list[int]:
        cached = getattr(self, "_int_pos_cache", None)
        if cached is None:
            cached = [i for i, key in enumerate(self.state_keys) if self.aliases.get(key, key) in self.int_state]
            self._int_pos_cache = cached
        return cached

    def _signature(self) -> tuple:
        sig = getattr(self, "_sig_cache", None)
        if sig is None:
            sig = (self.total, tuple((e.name, e.offset, e.numel, e.nhwc) for e in self.entries))
            self._sig_cache = sig
        return sig

    def same_layout(self, other: ParameterArena) -> bool:
        return other is self or self._signature() == other._signature()

    def range_of(self, names: Iterable[str]) -> list[tuple[int, int]]:
        """Merged (start, end) element ranges covering the given entries — the arena form of a layer subset."""
        spans = sorted((self.by_name[n].offset, _round_up(self.by_name[n].end, ALIGN)) for n in names if n in self.by_name)
        merged: list[tuple[int, int]] = []
        for start, end in spans:
            if merged and start <= merged[-1][1]:
                merged[-1] = (merged[-1][0], max(end, merged[-1][1]))
            else:
                merged.append((start, end))
        return merged


class TrainableRegionLayout:
    """Layout of a companion region restricted to the trainable parameters (in ``parameters()`` order): the shape
    SCAFFOLD control variates travel in.  Quacks like a ``ParameterArena`` for the aggregation fast path."""

    def __init__(self, arena: ParameterArena) -> None:
        self.arena = arena
        self.state_keys = [
            name for name, p in arena.module.named_parameters() if p.requires_grad and name in arena.by_name
        ]
        self.int_state: dict[str, torch.Tensor] = {}
        self.total = arena.trainable_padded
        self._views_cache: dict[tuple[int, int], NDArrays] = {}

    def ndarrays(self, names: Iterable[str] | None = None, region: torch.Tensor | None = None) -> NDArrays:
        assert region is not None
        if names is not None:
            return NDArrays([self.arena.view(name, region) for name in names])
        key = (region.data_ptr(), region.numel())
        cached = self._views_cache.get(key)
        if cached is None:  # the per-tensor views of a region are built once (a model's worth of slicing per call otherwise)
            if len(self._views_cache) > 8:
                self._views_cache.pop(next(iter(self._views_cache)))
            cached = self._views_cache[key] = NDArrays([self.arena.view(name, region) for name in self.state_keys])
            cached.flat, cached.layout = region[: self.total], self
        return NDArrays(cached, flat=cached.flat, layout=self)

    def same_layout(self, other: object) -> bool:
        return isinstance(other, TrainableRegionLayout) and self.arena.same_layout(other.arena)


def _own_version_alias(view: torch.Tensor) -> torch.Tensor:
    """The same memory as ``view`` behind a tensor with its OWN autograd version counter.  Views of one base share the
    base's counter: BatchNorm's in-place running-statistics update of one layer would then invalidate what another
    layer saved for its backward ("modified by an inplace operation") -- always the case when the arena is carved out
    of a single symmetric-memory segment, where every region of every arena is a view of the same base tensor."""
    alias = torch.empty(0, dtype=view.dtype, device=view.device)
    alias.data = view
    return alias


def _as_tensor(arr: object, device: torch.device) -> torch.Tensor:
    if isinstance(arr, torch.Tensor):
        return arr.to(device, non_blocking=True) if arr.device != device else arr
    import numpy as np

    return torch.from_numpy(np.asarray(arr, order="C").copy()).to(device, non_blocking=True)


def arena_of(module: nn.Module) -> ParameterArena | None:
    """The arena a module's state lives in, if any.  Kept in a weak side-table (not on the module) so that
    ``copy.deepcopy(model)`` / ``torch.save(model)`` see a plain module with ordinary tensors."""
    return _ARENAS.get(module)


def attach_arena(
    module: nn.Module,
    device: torch.device | str | None = None,
    allocator: Allocator | None = None,
    with_grad: bool = True,
    channels_last: bool = False,
) -> ParameterArena:
    existing = arena_of(module)
    if existing is not None and existing.module is module:
        return existing
    return ParameterArena(module, device=device, allocator=allocator, with_grad=with_grad, channels_last=channels_last)
