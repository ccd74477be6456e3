"""``PayloadSpec`` (parallel/spmd.py): how a client payload is described to the other ranks -- which leading entries are
one arena block, whether a second arena-shaped block is packed behind it (SCAFFOLD variates), whether the whole list is
a named subset of an arena (partial exchange), and what survives slicing by the packers."""

from __future__ import annotations

import numpy as np
import torch
from torch import nn

from fl4health_b200.common.typing import NDArrays, ndarrays_to_parameters, parameters_to_ndarrays
from fl4health_b200.parallel.arena import TrainableRegionLayout, attach_arena
from fl4health_b200.parallel.spmd import PayloadSpec, _slice_spec
from fl4health_b200.parameter_exchange.parameter_packer import (
    ParameterPackerAdaptiveConstraint,
    ParameterPackerWithControlVariates,
    ParameterPackerWithLayerNames,
)


def _net() -> nn.Module:
    torch.manual_seed(0)
    return nn.Sequential(nn.Conv2d(3, 4, 3), nn.BatchNorm2d(4), nn.Flatten(), nn.Linear(4, 2))


def test_whole_arena_payload_is_one_block() -> None:
    arena = attach_arena(_net())
    payload = arena.ndarrays()
    spec = PayloadSpec.of(payload)
    assert spec.is_arena and spec.flat_numel == arena.flat.numel() and spec.main_len == len(payload)
    assert all(inline is None for _, _, inline in spec.entries)  # counters included: they reduce on the device
    assert _slice_spec(spec, None, None, len(payload)) is spec


def test_scalar_packed_behind_the_weights_keeps_the_block() -> None:
    arena = attach_arena(_net())
    packer = ParameterPackerAdaptiveConstraint()
    packed = packer.pack_parameters(arena.ndarrays(), 0.25)
    assert packed.int_flat is arena.int_flat  # the packed counters travel with the packed list
    spec = PayloadSpec.of(packed)
    assert not spec.is_arena and spec.flat_numel == arena.flat.numel() and spec.main_len == len(packed) - 1
    assert float(spec.entries[-1][2]) == 0.25  # the scalar rides along by value
    weights_spec = _slice_spec(spec, None, -1, len(packed))
    assert weights_spec.is_arena and weights_spec.flat_numel == arena.flat.numel()
    # ... and through a Parameters round trip the weights slice is still arena-tagged
    weights, mu = packer.unpack_parameters(parameters_to_ndarrays(ndarrays_to_parameters(packed)))
    assert mu == 0.25 and weights.flat is arena.flat and weights.int_flat is arena.int_flat


def test_second_arena_block_for_control_variates() -> None:
    net = _net()
    arena = attach_arena(net)
    layout = TrainableRegionLayout(arena)
    variates = layout.ndarrays(region=arena.companion("delta_c", trainable_only=True))
    packer = ParameterPackerWithControlVariates(len(arena.state_keys))
    packed = packer.pack_parameters(arena.ndarrays(), variates)
    spec = PayloadSpec.of(packed)
    assert (spec.main_len, spec.aux_len) == (len(arena.state_keys), len(layout.state_keys))
    assert spec.flat_numel == arena.flat.numel() and spec.aux_numel == variates.flat.numel()
    split = packer.size_of_model_params
    first, second = _slice_spec(spec, None, split, len(packed)), _slice_spec(spec, split, None, len(packed))
    assert first.is_arena and first.flat_numel == spec.flat_numel and first.aux_numel is None
    assert second.is_arena and second.flat_numel == spec.aux_numel and second.main_len == spec.aux_len
    restored = parameters_to_ndarrays(ndarrays_to_parameters(packed))
    weights, variates_back = packer.unpack_parameters(restored)
    assert weights.flat is arena.flat and variates_back.flat.data_ptr() == variates.flat.data_ptr() and variates_back.layout is layout


def test_named_subset_of_an_arena() -> None:
    arena = attach_arena(_net())
    names = ["0.weight", "1.running_mean", "1.num_batches_tracked", "0.bias"]
    subset = arena.ndarrays(names)
    assert subset.flat is None and subset.subset_flat is arena.flat and subset.subset_names == tuple(names)
    assert arena.ndarrays(names)[0].data_ptr() == subset[0].data_ptr()  # (cached view list)
    spec = PayloadSpec.of(subset)
    assert spec.subset_numel == arena.flat.numel() and spec.subset_names == tuple(names) and not spec.is_arena
    ranges, int_positions = arena.subset_plan(tuple(names))
    assert int_positions == [2] and all(0 <= a < b <= arena.flat.numel() for a, b in ranges)
    covered = sum(b - a for a, b in ranges)
    assert covered >= arena.view("0.weight").numel() + arena.view("0.bias").numel() + 4
    # a subset with something packed behind it (dynamic layer exchange packs the names) is NOT described as a subset
    with_names = ParameterPackerWithLayerNames().pack_parameters(subset, names)
    assert PayloadSpec.of(with_names).subset_numel is None
    # pulling a result that is an arena-shaped buffer + names copies exactly those entries
    other = attach_arena(_net())
    result = torch.full_like(other.flat, 7.0)
    incoming = NDArrays([other.view(n, result) if n in other.by_name else torch.tensor(5) for n in names])
    incoming.subset_flat, incoming.subset_layout, incoming.subset_names = result, other, tuple(names)
    untouched = arena.view("3.weight").clone()
    arena.load_ndarrays(incoming, names)
    assert float(arena.view("0.weight").min()) == 7.0 and float(arena.view("1.running_mean").max()) == 7.0
    assert int(arena.int_state["1.num_batches_tracked"]) == 5 and torch.equal(arena.view("3.weight"), untouched)


def test_plain_lists_are_described_entry_by_entry() -> None:
    spec = PayloadSpec.of(NDArrays([torch.zeros(3, 2), np.array(["a", "b"]), torch.tensor(1.5)]))
    assert spec.flat_numel is None and spec.subset_numel is None and not spec.is_arena
    shapes = [shape for shape, _, _ in spec.entries]
    assert shapes == [(3, 2), (2,), ()] and spec.entries[0][2] is None and list(spec.entries[1][2]) == ["a", "b"]
    assert float(spec.entries[2][2]) == 1.5
