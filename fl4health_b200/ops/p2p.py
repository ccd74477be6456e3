"""Peer-memory (NVLink/NVSwitch) fused collectives: Python side of ``csrc/p2p_collectives.cu``.

``FusedCollectives`` owns this rank's *symmetric memory*: device allocations made with plain ``cudaMalloc`` (so they
can be exported with CUDA IPC), whose handles are exchanged once over the ``torch.distributed`` process group and
opened on every peer.  The result is a table ``peer_ptr[slot][rank]`` that the kernels dereference directly — P2P
loads/stores over NVLink issued from inside the kernel.  NCCL is only used to ship the 64-byte IPC handles.

* ``allocator``   — plug into ``ParameterArena(allocator=...)`` so a client's flat arena lives in symmetric memory
  (its contribution is then read in place by the peers: no staging copy);
* ``aggregate``   — ``agg_fused``: reduce-scatter + strategy epilogue + all-gather in one kernel;
* ``broadcast``   — ``bcast_fused``: scatter/all-gather from the root + receiver-side unpack in one kernel.
"""

from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from logging import INFO
from typing import Any

import torch

from fl4health_b200.common.logger import log
from fl4health_b200.ops import _lib

MAX_RANKS = 16


class _PeerArgs(ctypes.Structure):
    _fields_ = [
        ("contrib", ctypes.c_void_p * MAX_RANKS),
        ("result", ctypes.c_void_p * MAX_RANKS),
        ("flags", ctypes.c_void_p * MAX_RANKS),
        ("coef", ctypes.c_float * MAX_RANKS),
        ("rank", ctypes.c_int),
        ("world", ctypes.c_int),
    ]


class _RawBuffer:
    """A cudaMalloc'ed region exposed to torch through ``__cuda_array_interface__`` (zero-copy)."""

    def __init__(self, ptr: int, nbytes: int) -> None:
        self.ptr, self.nbytes = ptr, nbytes
        self.__cuda_array_interface__ = {
            "shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None,
        }


@dataclass
class _Slot:
    index: int
    ptr: int
    nbytes: int
    tensor: torch.Tensor  # uint8 view over the whole allocation (keeps the raw buffer alive)
    peer_ptrs: list[int] | None = None  # base address of the same slot on every rank (own address for own rank)
    opened: list[int] = field(default_factory=list)


class FusedCollectives:
    has_multicast = False  # NVLS multimem path is not wired yet; peers are reached with P2P loads/stores

    def __init__(self, ctx: Any) -> None:
        lib = _lib.load(True)
        assert lib is not None
        self.lib = lib
        self.ctx = ctx
        self.rank, self.world = ctx.rank, ctx.world_size
        if self.world > 8:
            raise RuntimeError("fused collectives support up to 8 ranks (one NVSwitch box)")
        device_index = ctx.device.index
        for peer in range(torch.cuda.device_count()):
            if peer != device_index and not lib.fl4h_can_access_peer(device_index, peer):
                raise RuntimeError(f"GPU {device_index} cannot access peer {peer}")
        self.slots: list[_Slot] = []
        self.epoch = 0
        self._result_by_numel: dict[int, tuple[_Slot, torch.Tensor]] = {}
        self._flags = self._alloc_slot(2 * MAX_RANKS * 4)
        self._exchange(self._flags)
        ctx.barrier()
        log(INFO, f"[fused] symmetric memory ready on rank {self.rank}/{self.world}")

    # -- symmetric allocation ---------------------------------------------------------------------------------
    def _alloc_slot(self, nbytes: int) -> _Slot:
        nbytes = (nbytes + 511) // 512 * 512
        ptr = ctypes.c_void_p()
        _lib.check(self.lib.fl4h_ipc_alloc(ctypes.c_size_t(nbytes), ctypes.byref(ptr)), "fl4h_ipc_alloc")
        raw = _RawBuffer(int(ptr.value), nbytes)
        tensor = torch.as_tensor(raw, device=self.ctx.device)
        tensor._fl4h_raw = raw  # type: ignore[attr-defined]
        slot = _Slot(len(self.slots), int(ptr.value), nbytes, tensor)
        self.slots.append(slot)
        return slot

    def allocator(self, numel: int, dtype: torch.dtype, device: torch.device) -> torch.Tensor:
        """Arena allocator: symmetric, zero-initialised.  Peers learn the address lazily at first collective use."""
        itemsize = torch.empty((), dtype=dtype).element_size()
        slot = self._alloc_slot(numel * itemsize)
        return slot.tensor[: numel * itemsize].view(dtype)

    def _slot_of(self, tensor: torch.Tensor) -> _Slot | None:
        ptr = tensor.data_ptr()
        for slot in self.slots:
            if slot.ptr <= ptr < slot.ptr + slot.nbytes:
                return slot
        return None

    def owns(self, tensor: torch.Tensor) -> bool:
        return tensor.is_cuda and self._slot_of(tensor) is not None

    def _exchange(self, slot: _Slot) -> None:
        """Collective: every rank contributes the IPC handle of its slot with this index."""
        if slot.peer_ptrs is not None:
            return
        handle = (ctypes.c_char * 64)()
        _lib.check(self.lib.fl4h_ipc_get_handle(ctypes.c_void_p(slot.ptr), handle), "fl4h_ipc_get_handle")
        gathered = self.ctx.all_gather_object((slot.index, slot.nbytes, bytes(handle)))
        peer_ptrs = []
        for rank, (index, nbytes, raw) in enumerate(gathered):
            if index != slot.index or nbytes != slot.nbytes:
                raise RuntimeError(
                    f"symmetric allocation order diverged between ranks (slot {slot.index}/{slot.nbytes} vs "
                    f"rank {rank}: {index}/{nbytes})"
                )
            if rank == self.rank:
                peer_ptrs.append(slot.ptr)
                continue
            opened = ctypes.c_void_p()
            buf = ctypes.create_string_buffer(raw, 64)
            _lib.check(self.lib.fl4h_ipc_open_handle(buf, ctypes.byref(opened)), "fl4h_ipc_open_handle")
            peer_ptrs.append(int(opened.value))
            slot.opened.append(int(opened.value))
        slot.peer_ptrs = peer_ptrs

    def result_buffer(self, numel: int) -> torch.Tensor:
        entry = self._result_by_numel.get(numel)
        if entry is None:
            slot = self._alloc_slot(numel * 4)
            entry = (slot, slot.tensor[: numel * 4].view(torch.float32))  # ONE tensor object: view caches key on it
            self._result_by_numel[numel] = entry
        return entry[1]

    # -- kernels ----------------------------------------------------------------------------------------------
    def _peer_args(self, contrib: torch.Tensor, result: torch.Tensor, coef_by_rank: list[float]) -> _PeerArgs:
        c_slot, r_slot = self._slot_of(contrib), self._slot_of(result)
        assert c_slot is not None and r_slot is not None, "buffers must live in symmetric memory"
        self._exchange(c_slot)
        self._exchange(r_slot)
        assert c_slot.peer_ptrs is not None and r_slot.peer_ptrs is not None and self._flags.peer_ptrs is not None
        c_off, r_off = contrib.data_ptr() - c_slot.ptr, result.data_ptr() - r_slot.ptr
        args = _PeerArgs()
        for rank in range(self.world):
            args.contrib[rank] = c_slot.peer_ptrs[rank] + c_off
            args.result[rank] = r_slot.peer_ptrs[rank] + r_off
            args.flags[rank] = self._flags.peer_ptrs[rank]
            args.coef[rank] = float(coef_by_rank[rank])
        args.rank, args.world = self.rank, self.world
        return args

    def aggregate(
        self, local: torch.Tensor, coef_by_rank: list[float], out: torch.Tensor | None = None,
        epilogue: dict[str, Any] | None = None,
    ) -> torch.Tensor:
        """``result = epilogue(sum_r coef[r] * flat_r)`` on every rank, one kernel (see module docstring)."""
        numel = local.numel()
        assert numel % 4 == 0 and local.dtype == torch.float32
        result = out if (out is not None and self.owns(out)) else self.result_buffer(numel)
        args = self._peer_args(local, result, coef_by_rank)
        epi = epilogue or {}
        self.epoch += 1
        err = self.lib.fl4h_agg_fused(
            ctypes.byref(args), _lib.ptr(epi.get("current")), _lib.ptr(epi.get("m")), _lib.ptr(epi.get("v")),
            ctypes.c_int(int(epi.get("mode", 0))), ctypes.c_float(epi.get("eta", 0.0)),
            ctypes.c_float(epi.get("beta1", 0.0)), ctypes.c_float(epi.get("beta2", 0.0)),
            ctypes.c_float(epi.get("tau", 0.0)), ctypes.c_float(epi.get("server_lr", 1.0)),
            ctypes.c_float(epi.get("momentum", 0.0)), ctypes.c_int64(numel), ctypes.c_uint32(self.epoch),
            _lib.stream_ptr(self.ctx.device),
        )
        _lib.check(err, "fl4h_agg_fused")
        _lib.count_launches(1)
        if out is not None and result.data_ptr() != out.data_ptr():
            out.copy_(result)
            return out
        return result

    def broadcast(
        self, src: torch.Tensor, root: int, w: torch.Tensor | None = None, anchor: torch.Tensor | None = None,
        shadow: torch.Tensor | None = None, c_server: torch.Tensor | None = None, c_local: torch.Tensor | None = None,
        cv_out: torch.Tensor | None = None,
    ) -> torch.Tensor:
        """Root's ``src`` (symmetric) -> every rank's result buffer + fused receiver-side unpack; returns the landed
        global buffer."""
        numel = src.numel()
        assert numel % 4 == 0 and src.dtype == torch.float32
        result = self.result_buffer(numel)
        args = self._peer_args(src, result, [0.0] * self.world)
        self.epoch += 1
        err = self.lib.fl4h_bcast_fused(
            ctypes.byref(args), ctypes.c_int(root), _lib.ptr(w), _lib.ptr(anchor), _lib.ptr(shadow),
            _lib.ptr(c_server), _lib.ptr(c_local), _lib.ptr(cv_out), ctypes.c_int64(numel),
            ctypes.c_uint32(self.epoch), _lib.stream_ptr(self.ctx.device),
        )
        _lib.check(err, "fl4h_bcast_fused")
        _lib.count_launches(1)
        return result

    def close(self) -> None:
        torch.cuda.synchronize(self.ctx.device)
        self.ctx.barrier()
        for slot in self.slots:
            for opened in slot.opened:
                self.lib.fl4h_ipc_close_handle(ctypes.c_void_p(opened))
            slot.opened = []
        # allocations are released with the process: tensors handed to arenas may still be referenced.
