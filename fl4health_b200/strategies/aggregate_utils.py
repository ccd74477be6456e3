"""Aggregation primitives.

Parity: ``fl4health/strategies/aggregate_utils.py:8-55`` (and Flower's ``aggregate`` / ``weighted_loss_avg``).  The
reference reduces per layer with ``functools.reduce(np.add, ...)`` over NumPy arrays on the server CPU.  Here:

* if every client's list is arena-backed (one flat buffer each, identical layout) the whole model is reduced by ONE
  launch of the K-way ``weighted_sum`` kernel over the flat buffers (fixed client order ⇒ bit-deterministic);
* otherwise each layer is reduced with device-side tensor ops (still no host round trip);
* string/object side arrays (layer names) are passed through untouched.
"""

from __future__ import annotations

from collections.abc import Sequence
from typing import Any

import numpy as np
import torch

from fl4health_b200.common.typing import NDArray, NDArrays, to_tensor
from fl4health_b200.ops import flat as flat_ops


def _common_flat(arrays: Sequence[NDArrays]) -> list[torch.Tensor] | None:
    """The clients' flat buffers if all lists are whole-arena views with one shared layout."""
    flats = []
    layout0 = None
    for nds in arrays:
        flat, layout = getattr(nds, "flat", None), getattr(nds, "layout", None)
        if flat is None or layout is None or len(nds) != len(layout.state_keys):
            return None
        if layout0 is None:
            layout0 = layout
        elif not layout.same_layout(layout0):
            return None
        flats.append(flat)
    if len({(f.device, f.dtype, f.numel()) for f in flats}) != 1:
        return None
    return flats


_RESULT_BUFFERS: dict[tuple, list[torch.Tensor]] = {}


def _result_buffer(layout: object, like: torch.Tensor) -> torch.Tensor:
    """Aggregation output buffers are recycled (two per layout, alternating) so that the per-layer view lists built
    over them stay cached and the allocator is not hit with a model-sized request every round.  Two buffers: the
    previous round's aggregate (= the server's current parameters) must stay intact while the next one is written."""
    key = (id(layout), like.numel(), like.device, like.dtype)
    ring = _RESULT_BUFFERS.setdefault(key, [])
    if len(ring) < 2:
        ring.append(torch.empty_like(like))
        return ring[-1]
    ring.append(ring.pop(0))
    return ring[-1]


def _merge_int_flats(int_flats: Sequence[torch.Tensor], coefficients: Sequence[float]) -> torch.Tensor:
    """``trunc(sum_k c_k * ints_k)`` over the clients' packed integer state: 2 tiny kernels per client, no host→device
    coefficient tensor (this runs right after the fit synchronisation point, i.e. with an idle GPU)."""
    acc = int_flats[0].to(torch.float64).mul_(float(coefficients[0]))
    for ints, coef in zip(int_flats[1:], coefficients[1:]):
        acc.add_(ints.to(torch.float64), alpha=float(coef))
    return acc.to(torch.int64)


def _scatter_int_views(out: NDArrays, layout: Any, merged: torch.Tensor) -> None:
    """Point the integer entries of ``out`` at slices of ``merged`` (one ``split`` + cached positions / shapes)."""
    plan = getattr(layout, "_int_scatter_plan", None)
    if plan is None:
        positions = [i for i, key in enumerate(layout.state_keys) if key in layout.int_state]
        states = [layout.int_state[layout.state_keys[i]] for i in positions]
        plan = (positions, [t.numel() for t in states], [t.shape for t in states], [t.dtype for t in states])
        layout._int_scatter_plan = plan
    positions, sizes, shapes, dtypes = plan
    out.int_flat = merged
    for idx, piece, shape, dtype in zip(positions, merged.split(sizes), shapes, dtypes):
        out[idx] = piece.view(shape) if dtype == merged.dtype else piece.view(shape).to(dtype)


def _is_meta_array(arr: NDArray) -> bool:
    return isinstance(arr, np.ndarray) and arr.dtype.kind in ("U", "S", "O")


def weighted_combine(arrays: Sequence[NDArrays], coefficients: Sequence[float]) -> NDArrays:
    """``sum_k coefficients[k] * arrays[k]`` layer-wise (or in one fused pass when arena-backed)."""
    assert len(arrays) == len(coefficients) and len(arrays) > 0
    if any(getattr(nds, "ctx", None) is not None for nds in arrays):
        return _spmd_weighted_combine(arrays, coefficients)
    flats = _common_flat(arrays)
    if flats is not None:
        layout = arrays[0].layout
        out_flat = _result_buffer(layout, flats[0])
        flat_ops.weighted_sum(out_flat, flats, coefficients)
        out = layout.ndarrays(region=out_flat)
        # integer state (e.g. num_batches_tracked) is not in the flat buffer: average it like the reference does
        int_flats = [getattr(nds.layout, "int_flat", None) for nds in arrays]
        if layout.int_state and all(f is not None for f in int_flats):
            _scatter_int_views(out, layout, _merge_int_flats(int_flats, coefficients))
        else:
            out.int_flat = None
            for idx, key in enumerate(layout.state_keys):
                if key in layout.int_state:
                    out[idx] = _combine_layer([nds[idx] for nds in arrays], coefficients)
        return out
    n_layers = len(arrays[0])
    assert all(len(nds) == n_layers for nds in arrays), "clients sent different numbers of arrays"
    return NDArrays([_combine_layer([nds[i] for nds in arrays], coefficients) for i in range(n_layers)])


def _spmd_weighted_combine(
    arrays: Sequence[NDArrays], coefficients: Sequence[float], epilogue: dict | None = None,
    out_flat: torch.Tensor | None = None,
) -> NDArrays:
    """SPMD form of ``weighted_combine``: each rank owns at most one entry of ``arrays`` (its own client's payload);
    the others are ``RemoteNDArrays`` placeholders.  One collective over the flat payload reduces them all."""
    from fl4health_b200.parallel.spmd import PayloadSpec

    ctx = next(nds.ctx for nds in arrays if getattr(nds, "ctx", None) is not None)
    if len({nds.rank for nds in arrays}) < len(arrays):  # some rank hosts several of these clients
        return _spmd_weighted_combine_multi(ctx, arrays, coefficients, epilogue, out_flat)
    coef_by_rank = [0.0] * ctx.world_size
    local: NDArrays | None = None
    spec: PayloadSpec | None = None
    for nds, coef in zip(arrays, coefficients):
        coef_by_rank[nds.rank] = float(coef)
        spec = nds.spec
        if nds.rank == ctx.rank:
            local = nds
    assert spec is not None
    subset = _subset_combine(ctx, arrays, coef_by_rank, local, spec) if epilogue is None else None
    if subset is not None:
        return subset
    arena_route = spec.is_arena
    if arena_route:
        layout = local.layout if local is not None else None
        local_flat = local.flat if local is not None else None
        if layout is not None:
            _LAYOUT_HINTS[spec.flat_numel] = layout
        if len(arrays) < ctx.world_size:
            # Partial participation (fraction / Poisson sampling): a rank whose client sat this round out contributes
            # weight 0 from its own arena, whose layout it knows from an earlier round — that also keeps the fused
            # peer-memory kernel's launch decision identical on every rank.  If some rank has never seen the layout,
            # everybody takes the packed route below (one agreement collective, only in this mode).
            hint = layout if layout is not None else _LAYOUT_HINTS.get(spec.flat_numel)
            if hint is not None and not hasattr(hint, "flat"):
                hint = None  # a companion-region layout has no buffer of its own to stand in with
            if ctx.all_reduce_max(0.0 if hint is not None else 1.0) > 0.0:
                arena_route = False
            elif layout is None:
                layout, local_flat = hint, hint.flat[: spec.flat_numel]
    if arena_route:
        if out_flat is None and local_flat is not None and layout is not None and ctx.fused is None:
            out_flat = _result_buffer(layout, local_flat)
        int_flat = getattr(layout, "int_flat", None) if layout is not None else None
        result_flat = ctx.weighted_sum_flat(local_flat, coef_by_rank, spec.flat_numel, out=out_flat, epilogue=epilogue,
                                            int_local=int_flat)
        if layout is None:
            raise RuntimeError("rank without a local payload cannot rebuild arena views; sample all ranks or use weight 0")
        out = layout.ndarrays(region=result_flat)
        int_idx = [i for i, key in enumerate(layout.state_keys) if key in layout.int_state]
        if int_idx:
            reduced = getattr(ctx, "last_int_reduced", None)
            if reduced is not None and reduced.numel() == sum(out[i].numel() for i in int_idx):
                ints = reduced  # the fused aggregate kernel already reduced the counters (same launch)
            else:
                if int_flat is not None and int_flat.numel() == sum(out[i].numel() for i in int_idx):
                    ints = int_flat.to(torch.float64) * coef_by_rank[ctx.rank]  # one tensor for all integer buffers
                else:
                    ints = torch.cat([out[i].reshape(-1).to(torch.float64) * coef_by_rank[ctx.rank] for i in int_idx])
                if ctx.world_size > 1:
                    import torch.distributed as dist

                    from fl4health_b200.utils import tracing

                    with tracing.phase("agg_int_buffers"):
                        dist.all_reduce(ints)
            _scatter_int_views(out, layout, ints.to(torch.int64))
        return out
    # non-arena payloads (or no agreed layout): pack the tensor entries into one temporary flat buffer, reduce, unpack
    assert epilogue is None, "server-optimizer epilogues need arena-backed payloads (seen by every rank) in SPMD mode"
    tensor_idx = [i for i, (_, _, inline) in enumerate(spec.entries) if inline is None]
    shapes = [spec.entries[i][0] for i in tensor_idx]
    sizes = [int(np.prod(s)) if len(s) else 1 for s in shapes]
    total = sum(sizes)
    packed = None
    padded = (total + 3) // 4 * 4  # the flat kernels work on 16-byte vectors
    if local is not None:
        packed = torch.zeros(padded, dtype=torch.float32, device=ctx.device)
        if tensor_idx:
            packed[:total] = torch.cat([to_tensor(local[i], ctx.device).reshape(-1).to(torch.float32) for i in tensor_idx])
    reduced = ctx.weighted_sum_flat(packed, coef_by_rank, padded)
    out = NDArrays([entry[2] for entry in spec.entries])
    cursor = 0
    for i, shape, size in zip(tensor_idx, shapes, sizes):
        out[i] = reduced[cursor : cursor + size].view(shape)
        cursor += size
    return out


def _subset_combine(ctx: Any, arrays: Sequence[NDArrays], coef_by_rank: list[float], local: NDArrays | None, spec: Any) -> NDArrays | None:
    """Partial exchange (FedPer, FedRep, FedBN ...): every payload is the same NAMED SUBSET of its rank's arena.  The
    whole arenas are reduced with the one fused launch of a full exchange (a fraction of a millisecond; what is not
    exchanged is averaged too and simply never read) and the result is handed out as views of the result buffer at the
    subset's offsets -- instead of concatenating, reducing and re-slicing a hundred tensors in Python every round."""
    if spec.subset_numel is None or local is None or len(arrays) < ctx.world_size:
        return None
    if any(nds.spec.subset_numel != spec.subset_numel or nds.spec.subset_names != spec.subset_names for nds in arrays):
        return None
    layout, local_flat = getattr(local, "subset_layout", None), getattr(local, "subset_flat", None)
    if layout is None or local_flat is None or not hasattr(layout, "subset_plan"):
        return None
    names = spec.subset_names
    int_flat = getattr(layout, "int_flat", None)
    out_flat = _result_buffer(layout, local_flat) if ctx.fused is None else None
    result_flat = ctx.weighted_sum_flat(local_flat, coef_by_rank, spec.subset_numel, out=out_flat, int_local=int_flat)
    cache = layout.__dict__.setdefault("_subset_result_views", {})
    key = (result_flat.data_ptr(), names)
    views = cache.get(key)
    if views is None:
        if len(cache) > 8:
            cache.pop(next(iter(cache)))
        views = cache[key] = [layout.view(name, result_flat) if layout.aliases.get(name, name) in layout.by_name else None for name in names]
    out = NDArrays(views)
    _, int_positions = layout.subset_plan(names)
    if int_positions:
        reduced = getattr(ctx, "last_int_reduced", None)
        all_ints = [key for key in layout.state_keys if key in layout.int_state]
        if reduced is None or reduced.numel() != sum(layout.int_state[k].numel() for k in all_ints):
            reduced = int_flat.to(torch.float64) * coef_by_rank[ctx.rank]
            if ctx.world_size > 1:
                import torch.distributed as dist

                dist.all_reduce(reduced)
            reduced = reduced.to(torch.int64)
        offsets, cursor = {}, 0
        for key in all_ints:
            offsets[key] = cursor
            cursor += layout.int_state[key].numel()
        for position in int_positions:
            key = layout.aliases.get(names[position], names[position])
            state = layout.int_state[key]
            out[position] = reduced[offsets[key]:offsets[key] + state.numel()].view(state.shape).to(state.dtype)
    out.subset_flat, out.subset_layout, out.subset_names = result_flat, layout, names
    return out


_LAYOUT_HINTS: dict[int, Any] = {}  # flat_numel -> arena layout last seen on this rank (ranks whose clients sat a round out)


def _spmd_weighted_combine_multi(
    ctx: Any, arrays: Sequence[NDArrays], coefficients: Sequence[float], epilogue: dict | None = None,
    out_flat: torch.Tensor | None = None,
) -> NDArrays:
    """``parallel/spmd_multi.py``: ranks host several clients.  Each rank folds ITS payloads into one partial sum with a
    streaming kernel, the partials are reduced with one all-reduce (cost independent of the number of clients), then the
    strategy epilogue runs.  Payloads that are not arena-shaped, or a rank that has never seen the layout, take the
    always-correct route: materialise every payload everywhere and combine locally."""

    spec = arrays[0].spec
    local = [(nds, float(c)) for nds, c in zip(arrays, coefficients) if nds.rank == ctx.rank and not getattr(nds, "remote", False)]
    layout = next((nds.layout for nds, _ in local if getattr(nds, "layout", None) is not None), None)
    arena_shaped = spec.is_arena and all(nds.spec.is_arena and nds.spec.flat_numel == spec.flat_numel for nds in arrays)
    if arena_shaped and layout is not None:
        _LAYOUT_HINTS[spec.flat_numel] = layout
    elif arena_shaped:
        layout = _LAYOUT_HINTS.get(spec.flat_numel)
    # the choice below must be the same on every rank: agree on it (a rank that does not know the layout cannot rebuild the
    # arena views, so everybody takes the packed route)
    if ctx.all_reduce_max(0.0 if (arena_shaped and layout is not None) else 1.0) > 0.0:
        assert epilogue is None, "server-optimizer epilogues need arena-backed payloads in SPMD mode"
        # packed route (side payloads such as SCAFFOLD variates, models without an arena): the tensor entries of every
        # local payload are concatenated, scaled and summed into one buffer, reduced with one all-reduce, and unpacked
        tensor_idx = [i for i, (_, _, inline) in enumerate(spec.entries) if inline is None]
        shapes = [spec.entries[i][0] for i in tensor_idx]
        sizes = [int(np.prod(shape)) if len(shape) else 1 for shape in shapes]
        total = sum(sizes)
        packed = torch.zeros((total + 3) // 4 * 4, dtype=torch.float32, device=ctx.device)  # 16-byte vectors in the flat kernels
        for nds, coef in local:
            if tensor_idx:
                packed[:total].add_(torch.cat([to_tensor(nds[i], ctx.device).reshape(-1).to(torch.float32) for i in tensor_idx]), alpha=coef)
        reduced = ctx.weighted_sum_flat(packed, [1.0] * ctx.world_size, packed.numel())
        out = NDArrays([entry[2] for entry in spec.entries])
        cursor = 0
        for i, shape, size in zip(tensor_idx, shapes, sizes):
            out[i] = reduced[cursor : cursor + size].view(shape)
            cursor += size
        return out
    numel = spec.flat_numel
    partial = torch.zeros(numel, dtype=torch.float32, device=ctx.device)
    if local:
        flat_ops.weighted_sum(partial, [nds.flat[:numel] for nds, _ in local], [c for _, c in local])
    result_flat = ctx.weighted_sum_flat(partial, [1.0] * ctx.world_size, numel, out=out_flat, epilogue=epilogue)
    out = layout.ndarrays(region=result_flat)
    if layout.int_state:
        positions = [i for i, key in enumerate(layout.state_keys) if key in layout.int_state]
        width = sum(layout.int_state[layout.state_keys[i]].numel() for i in positions)
        ints = torch.zeros(width, dtype=torch.float64, device=ctx.device)
        for nds, coef in local:
            packed = getattr(nds, "int_flat", None)
            if packed is None or packed.numel() != width:
                packed = torch.cat([to_tensor(nds[i], ctx.device).reshape(-1) for i in positions])
            ints.add_(packed.to(torch.float64), alpha=coef)
        if ctx.world_size > 1:
            import torch.distributed as dist

            dist.all_reduce(ints)
        _scatter_int_views(out, layout, ints.to(torch.int64))
    return out


def _combine_layer(layers: Sequence[NDArray], coefficients: Sequence[float]) -> NDArray:
    first = layers[0]
    if _is_meta_array(first):
        return first
    if all(isinstance(layer, np.ndarray) for layer in layers):
        acc = np.asarray(layers[0]) * coefficients[0]
        for layer, coef in zip(layers[1:], coefficients[1:]):
            acc = acc + np.asarray(layer) * coef
        return acc
    device = next((layer.device for layer in layers if isinstance(layer, torch.Tensor)), None)
    tensors = [to_tensor(layer, device) for layer in layers]
    work_dtype = tensors[0].dtype if tensors[0].is_floating_point() else torch.float64
    acc = tensors[0].to(work_dtype) * coefficients[0]
    for tensor, coef in zip(tensors[1:], coefficients[1:]):
        acc = acc + tensor.to(work_dtype) * coef
    return acc if tensors[0].is_floating_point() else acc.to(tensors[0].dtype)


def aggregate_results(results: list[tuple[NDArrays, int]], weighted: bool = True) -> NDArrays:
    """Weighted (by sample count) or uniform average of client array lists."""
    arrays = [nds for nds, _ in results]
    if weighted:
        total = sum(n for _, n in results)
        coefficients = [n / total for _, n in results]
    else:
        coefficients = [1.0 / len(results)] * len(results)
    return weighted_combine(arrays, coefficients)


def aggregate_losses(results: list[tuple[int, float]], weighted: bool = True) -> float:
    """Average client losses; sorted first so the float sum is order-independent."""
    ordered = sorted(results, key=lambda item: item[1])
    if weighted:
        total = sum(n for n, _ in ordered)
        return sum(n * loss for n, loss in ordered) / total
    return sum(loss for _, loss in ordered) / len(ordered)


def weighted_loss_avg(results: list[tuple[int, float]]) -> float:
    total = sum(n for n, _ in results)
    return sum(n * loss for n, loss in results) / total
