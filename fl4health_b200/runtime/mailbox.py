"""``ShmMailbox``: all-gather of small float64 records between the ranks of one node through shared memory (native
implementation in ``csrc/shm_mailbox.cpp``).  Used by ``parallel/spmd.py`` for the per-round result metadata, so the
exchange neither launches a collective nor synchronises a CUDA stream."""

from __future__ import annotations

import ctypes
import os
from typing import Any

import numpy as np

from fl4health_b200.runtime import build as _build

_LIB: Any = None


def load_runtime() -> Any:
    """The host runtime library, built on first use when a compiler is present; ``None`` when unavailable."""
    global _LIB
    if _LIB is None:
        try:
            path = _build.build()
            lib = ctypes.CDLL(str(path))
        except (RuntimeError, OSError):
            _LIB = False
            return None
        lib.fl4h_mbox_open.restype = ctypes.c_void_p
        lib.fl4h_mbox_open.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        lib.fl4h_mbox_unlink.argtypes = [ctypes.c_char_p]
        lib.fl4h_mbox_post.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_int]
        lib.fl4h_mbox_gather.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double]
        lib.fl4h_mbox_close.argtypes = [ctypes.c_void_p]
        _LIB = lib
    return _LIB or None


class ShmMailbox:
    """One mailbox per process group.  ``name`` must be the same on every rank; rank 0 creates the segment first (the
    caller orders creation / opening / unlinking with its own rendezvous, see ``SpmdContext._open_mailbox``)."""

    def __init__(self, name: str, world: int, rank: int, capacity: int = 256, create: bool = False) -> None:
        lib = load_runtime()
        if lib is None:
            raise RuntimeError("host runtime library unavailable")
        self.lib, self.name, self.world, self.rank, self.capacity = lib, name, world, rank, capacity
        self.handle = lib.fl4h_mbox_open(name.encode(), world, rank, capacity, int(create))
        if not self.handle:
            raise RuntimeError(f"could not {'create' if create else 'open'} shared-memory mailbox {name!r}")
        self.seq = 0
        self._out = np.empty((world, capacity), dtype=np.float64)
        self._lens = np.zeros(world, dtype=np.int32)

    def unlink(self) -> None:
        self.lib.fl4h_mbox_unlink(self.name.encode())

    def all_gather(self, values: np.ndarray | list[float], timeout: float = 600.0) -> list[np.ndarray]:
        """Every rank's record (collective: all ranks must call it the same number of times)."""
        record = np.ascontiguousarray(values, dtype=np.float64)
        if record.size > self.capacity:
            raise ValueError(f"record of {record.size} doubles exceeds the mailbox capacity {self.capacity}")
        self.seq += 1
        if self.lib.fl4h_mbox_post(self.handle, self.seq, record.ctypes.data, int(record.size)) != 0:
            raise RuntimeError("mailbox post failed")
        rc = self.lib.fl4h_mbox_gather(self.handle, self.seq, self._out.ctypes.data, self._lens.ctypes.data, float(timeout))
        if rc != 0:
            raise TimeoutError(f"mailbox gather timed out after {timeout}s (a peer rank died or left the round loop)")
        return [self._out[r, : self._lens[r]].copy() for r in range(self.world)]

    def close(self) -> None:
        if self.handle:
            self.lib.fl4h_mbox_close(self.handle)
            self.handle = None


def default_name() -> str:
    return f"/fl4h_mbox_{os.getuid()}_{os.getpid()}"
