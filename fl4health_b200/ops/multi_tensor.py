"""Multi-tensor ("table") optimizer step: per-parameter gradients (bf16 or fp32, wherever autograd put them) are
consumed through a pointer table carried in the kernel parameter block; masters, moments, FedProx anchor, SCAFFOLD
correction and the bf16 compute shadow are flat arena regions.  Kernel: ``csrc/mt_optim.cu``.  The CPU path runs the
same per-slice reference math as ``ops.flat`` (and is the numerics oracle for the GPU test).
"""

from __future__ import annotations

import ctypes
from dataclasses import dataclass

import torch

from fl4health_b200.ops import _lib
from fl4health_b200.ops import flat as F

MT_MAX = 512
MT_CHUNK = 4096


class _MtTable(ctypes.Structure):
    _fields_ = [
        ("grad", ctypes.c_void_p * MT_MAX),
        ("offset", ctypes.c_int64 * MT_MAX),
        ("numel", ctypes.c_int32 * MT_MAX),
        ("chunk_prefix", ctypes.c_int32 * (MT_MAX + 1)),
        ("grad_bf16", ctypes.c_uint8 * MT_MAX),
        ("count", ctypes.c_int32),
    ]


@dataclass
class TableEntry:
    grad: torch.Tensor  # physical order == arena physical order (checked by the caller)
    offset: int
    numel: int


def _launch_tables(entries: list[TableEntry]) -> list[_MtTable]:
    tables = []
    for start in range(0, len(entries), MT_MAX):
        chunk = entries[start : start + MT_MAX]
        table = _MtTable()
        prefix = 0
        for i, entry in enumerate(chunk):
            table.grad[i] = entry.grad.data_ptr()
            table.offset[i] = entry.offset
            table.numel[i] = entry.numel
            table.grad_bf16[i] = 1 if entry.grad.dtype == torch.bfloat16 else 0
            table.chunk_prefix[i] = prefix
            prefix += (entry.numel + MT_CHUNK - 1) // MT_CHUNK
        table.chunk_prefix[len(chunk)] = prefix
        table.count = len(chunk)
        tables.append(table)
    return tables


def mt_step(
    entries: list[TableEntry],
    adam: bool,
    w: torch.Tensor,
    m1: torch.Tensor | None,
    m2: torch.Tensor | None,
    hp: torch.Tensor,
    anchor: torch.Tensor | None = None,
    cv: torch.Tensor | None = None,
    shadow: torch.Tensor | None = None,
    decoupled: bool = True,
) -> None:
    """One optimizer step over every (gradient, arena slice) pair in ``entries``.  SGD clears the first-step flag and
    Adam ticks its step counter exactly once per call, as the flat kernels do."""
    if not entries:
        return
    for entry in entries:
        assert entry.grad.dtype in (torch.float32, torch.bfloat16), f"unsupported gradient dtype {entry.grad.dtype}"
    if F._use_kernel(w):
        lib = _lib.load(True)
        if ctypes.sizeof(_MtTable) != lib.fl4h_mt_table_size():
            raise RuntimeError("MtTable ABI mismatch between multi_tensor.py and mt_optim.cu")
        stream = _lib.stream_ptr(w.device)
        for i, table in enumerate(_launch_tables(entries)):
            err = lib.fl4h_mt_step(
                ctypes.byref(table), ctypes.c_int(1 if adam else 0), _lib.ptr(w), _lib.ptr(m1), _lib.ptr(m2),
                _lib.ptr(anchor), _lib.ptr(cv), _lib.ptr(shadow), _lib.ptr(hp), ctypes.c_int(1 if decoupled else 0),
                ctypes.c_int(1 if i == 0 else 0), stream,
            )
            _lib.check(err, "fl4h_mt_step")
            _lib.count_launches(2 if (adam and i == 0) else 1)
        if not adam:
            _lib.check(lib.fl4h_mt_clear_first(_lib.ptr(hp), stream), "fl4h_mt_clear_first")
            _lib.count_launches(1)
        return
    mt_step_reference(entries, adam, w, m1, m2, hp, anchor, cv, shadow, decoupled)


def mt_step_reference(entries, adam, w, m1, m2, hp, anchor=None, cv=None, shadow=None, decoupled=True) -> None:  # noqa: ANN001
    first = float(hp[F.HP_FIRST])
    step = float(hp[F.HP_STEP])
    for entry in entries:
        lo, hi = entry.offset, entry.offset + entry.numel
        g = entry.grad.reshape(-1) if entry.grad.is_contiguous() else _physical_flat(entry.grad)
        sl = lambda t: None if t is None else t[lo:hi]  # noqa: E731
        if adam:
            hp[F.HP_STEP] = step  # every tensor sees the same (pre-tick) counter
            F.adamw_step_reference(w[lo:hi], g, m1[lo:hi], m2[lo:hi], hp, sl(anchor), sl(shadow), decoupled)
        else:
            hp[F.HP_FIRST] = first
            F.sgd_step_reference(w[lo:hi], g, sl(m1), hp, sl(anchor), sl(cv), sl(shadow))
    if adam:
        hp[F.HP_STEP] = step + 1.0
    else:
        hp[F.HP_FIRST] = 0.0


def _physical_flat(t: torch.Tensor) -> torch.Tensor:
    """1-D view/copy of ``t`` in its physical (stride) order; for dense permuted tensors (channels-last) no copy."""
    order = sorted(range(t.dim()), key=lambda d: -t.stride(d))
    return t.permute(order).reshape(-1)
