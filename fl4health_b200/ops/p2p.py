"""NVLink / NVSwitch fused collectives: Python side of ``csrc/p2p_collectives.cu`` + ``csrc/symm_mem.cu``.

``FusedCollectives`` owns this rank's *symmetric memory*.  Memory comes in **segments**: one CUDA-VMM physical
allocation per rank (``cuMemCreate``), created collectively and

* mapped into every peer process (unicast P2P addresses, ``peer_ptrs[rank]``): the file descriptors of the exported
  handles travel between the rank processes over AF_UNIX sockets (``SCM_RIGHTS``), NCCL is not involved;
* bound to an NVSwitch **multicast object** (``cuMulticastCreate/AddDevice/BindMem``) whose mapping ``mc_ptr`` is what
  the kernels hand to ``multimem.ld_reduce`` (in-switch sum over all ranks' copies) and ``multimem.st`` (in-switch
  replication into all copies) — NVLS.

Buffers (parameter arenas, result / staging buffers, integer model buffers, signal flags) are bump-allocated inside
segments, so every buffer has the same offset on every rank.

* ``allocator``   — plug into ``ParameterArena(allocator=...)`` so a client's flat arena lives in symmetric memory
  (its contribution is then reduced in place: no staging copy when the FedAvg weights are uniform);
* ``aggregate``   — reduce-scatter + strategy epilogue + all-gather in ONE kernel (``agg_nvls`` / ``agg_fused``),
  integer buffers (BatchNorm counters) reduced by the same launch;
* ``broadcast``   — root -> every rank (``multimem.st``) + receiver-side unpack (w, FedProx anchor, bf16 shadow,
  SCAFFOLD ``c - c_i``) in ONE kernel (``bcast_nvls`` / ``bcast_fused``).

``FL4H_NVLS=0`` selects the fixed-order P2P kernels (bit-deterministic reduction order), ``FL4H_NVLS=1`` the multimem
kernels; by default the path that moves fewer bytes for the world size is taken (P2P below 4 ranks, NVLS from 4).
"""

from __future__ import annotations

import ctypes
import os
import socket
import tempfile
import time
from dataclasses import dataclass, field
from logging import INFO, WARNING
from typing import Any

import torch

from fl4health_b200.common.logger import log
from fl4health_b200.ops import _lib

MAX_RANKS = 16
_ALIGN = 512  # bytes; every buffer starts on a 512-byte boundary inside its segment


class _PeerArgs(ctypes.Structure):
    _fields_ = [
        ("contrib", ctypes.c_void_p * MAX_RANKS),
        ("result", ctypes.c_void_p * MAX_RANKS),
        ("flags", ctypes.c_void_p * MAX_RANKS),
        ("coef", ctypes.c_float * MAX_RANKS),
        ("rank", ctypes.c_int),
        ("world", ctypes.c_int),
        ("mc_contrib", ctypes.c_void_p),
        ("mc_result", ctypes.c_void_p),
        ("stage", ctypes.c_void_p),
        ("ibuf", ctypes.c_void_p * MAX_RANKS),
        ("ibuf_out", ctypes.c_void_p),
        ("n_int", ctypes.c_int),
        ("use_nvls", ctypes.c_int),
        ("uniform_coef", ctypes.c_float),
    ]


class _RawBuffer:
    """A device region exposed to torch through ``__cuda_array_interface__`` (zero-copy)."""

    def __init__(self, ptr: int, nbytes: int) -> None:
        self.ptr, self.nbytes = ptr, nbytes
        self.__cuda_array_interface__ = {
            "shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None,
        }


@dataclass
class _Segment:
    index: int
    nbytes: int
    ptr: int  # this rank's unicast mapping
    handle: int
    fd: int  # exported handle, consumed by the exchange
    tensor: torch.Tensor  # uint8 view over the whole segment
    cursor: int = 0
    peer_ptrs: list[int] | None = None  # unicast mapping of every rank's copy (own ptr for own rank)
    peer_handles: list[int] = field(default_factory=list)
    mc_ptr: int = 0  # multicast mapping (0: not bound)
    mc_handle: int = 0


class _FdMesh:
    """Full mesh of AF_UNIX stream sockets between the rank processes of one node (for SCM_RIGHTS fd passing)."""

    def __init__(self, ctx: Any) -> None:
        self.rank, self.world = ctx.rank, ctx.world_size
        token = ctx.broadcast_object(f"{os.getpid()}_{time.time_ns()}", src=0)
        base = os.path.join(tempfile.gettempdir(), f"fl4h_symm_{token}")
        my_path = f"{base}_{self.rank}.sock"
        listener = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        listener.bind(my_path)
        listener.listen(self.world)
        ctx.barrier()  # every listener is up
        self.socks: dict[int, socket.socket] = {}
        for peer in range(self.rank):  # connect "down", accept "up": each pair gets exactly one connection
            s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            s.connect(f"{base}_{peer}.sock")
            s.sendall(self.rank.to_bytes(4, "little"))
            self.socks[peer] = s
        for _ in range(self.rank + 1, self.world):
            s, _addr = listener.accept()
            peer = int.from_bytes(self._recv_exact(s, 4), "little")
            self.socks[peer] = s
        ctx.barrier()
        listener.close()
        os.unlink(my_path)

    @staticmethod
    def _recv_exact(sock: socket.socket, n: int) -> bytes:
        buf = b""
        while len(buf) < n:
            chunk = sock.recv(n - len(buf))
            if not chunk:
                raise ConnectionError("peer closed the fd-exchange socket")
            buf += chunk
        return buf

    def all_to_all_fd(self, fd: int) -> dict[int, int]:
        """Send a duplicate of ``fd`` to every peer, receive one from every peer."""
        for peer in sorted(self.socks):
            socket.send_fds(self.socks[peer], [b"F"], [fd])
        received = {}
        for peer in sorted(self.socks):
            _msg, fds, _flags, _addr = socket.recv_fds(self.socks[peer], 1, 1)
            received[peer] = fds[0]
        return received

    def broadcast_fd(self, fd: int | None, src: int = 0) -> int | None:
        if self.rank == src:
            for peer in sorted(self.socks):
                socket.send_fds(self.socks[peer], [b"M"], [fd])
            return None
        _msg, fds, _flags, _addr = socket.recv_fds(self.socks[src], 1, 1)
        return fds[0]

    def close(self) -> None:
        for s in self.socks.values():
            s.close()
        self.socks = {}


class FusedCollectives:
    def __init__(self, ctx: Any) -> None:
        lib = _lib.load(True)
        assert lib is not None
        self.lib = lib
        self.ctx = ctx
        self.rank, self.world = ctx.rank, ctx.world_size
        if self.world > 8:
            raise RuntimeError("fused collectives support up to 8 ranks (one NVSwitch box)")
        self.device_index = ctx.device.index
        features = lib.fl4h_symm_features(self.device_index)
        if features & 3 != 3:
            raise RuntimeError(f"CUDA VMM with POSIX-fd handles is not available on this device (features={features})")
        for peer in range(torch.cuda.device_count()):
            if peer != self.device_index and not lib.fl4h_can_access_peer(self.device_index, peer):
                raise RuntimeError(f"GPU {self.device_index} cannot access peer {peer}")
        # FL4H_NVLS: 1 = multimem kernels, 0 = fixed-order P2P kernels, unset = by traffic: per direction and GPU the
        # NVLS aggregate moves payload*(1 + 1/K), the P2P one 2*payload*(K-1)/K  ->  P2P wins at K = 2 (measured 100 vs
        # 158 us for 44.7 MB), they tie at K = 3, NVLS wins from K = 4 (1.25 vs 1.5, 1.125 vs 1.75 at K = 8)
        policy = os.environ.get("FL4H_NVLS", "auto")
        want_nvls = policy == "1" or (policy != "0" and self.world >= 4)
        # every rank must take the same decision: multicast needs all devices to support it
        self.has_multicast = bool(want_nvls and self.world > 1
                                  and ctx.all_reduce_max(0.0 if features & 4 else 1.0) == 0.0)
        self.segment_bytes = int(os.environ.get("FL4H_SYMM_SEGMENT_MB", "256")) << 20
        self.mesh = _FdMesh(ctx) if self.world > 1 else None
        self.segments: list[_Segment] = []
        self.epoch = 0
        self._result_by_key: dict[tuple[str, int], torch.Tensor] = {}
        self._flags = self._carve(2 * MAX_RANKS * 4, torch.uint8)
        self._exchange(self._segment_of(self._flags))
        ctx.barrier()
        log(INFO, f"[fused] symmetric memory ready on rank {self.rank}/{self.world} (multicast={self.has_multicast})")

    # -- symmetric allocation ---------------------------------------------------------------------------------
    def _new_segment(self, min_bytes: int) -> _Segment:
        want = max(min_bytes, self.segment_bytes)
        rounded = ctypes.c_size_t()
        _lib.check(self.lib.fl4h_symm_round_size(self.device_index, self.world, int(self.has_multicast),
                                                 ctypes.c_size_t(want), ctypes.byref(rounded)), "fl4h_symm_round_size")
        handle, ptr, fd = ctypes.c_ulonglong(), ctypes.c_void_p(), ctypes.c_int()
        _lib.check(self.lib.fl4h_symm_create(self.device_index, rounded, ctypes.byref(handle), ctypes.byref(ptr),
                                             ctypes.byref(fd)), "fl4h_symm_create")
        raw = _RawBuffer(int(ptr.value), int(rounded.value))
        tensor = torch.as_tensor(raw, device=self.ctx.device)
        tensor._fl4h_raw = raw  # type: ignore[attr-defined]
        seg = _Segment(len(self.segments), int(rounded.value), int(ptr.value), int(handle.value), int(fd.value), tensor)
        self.segments.append(seg)
        return seg

    def _carve(self, nbytes: int, dtype: torch.dtype) -> torch.Tensor:
        nbytes = (nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
        seg = next((s for s in self.segments if s.nbytes - s.cursor >= nbytes), None)
        if seg is None:
            seg = self._new_segment(nbytes)
        view = seg.tensor[seg.cursor : seg.cursor + nbytes]
        seg.cursor += nbytes
        return view.view(dtype)

    def allocator(self, numel: int, dtype: torch.dtype, device: torch.device) -> torch.Tensor:
        """Arena allocator: symmetric, zero-initialised.  Peers map the segment lazily at first collective use (every
        rank allocates in the same order, so offsets agree)."""
        itemsize = torch.empty((), dtype=dtype).element_size()
        return self._carve(numel * itemsize, torch.uint8)[: numel * itemsize].view(dtype)

    def _segment_of(self, tensor: torch.Tensor) -> _Segment | None:
        ptr = tensor.data_ptr()
        for seg in self.segments:
            if seg.ptr <= ptr < seg.ptr + seg.nbytes:
                return seg
        return None

    def owns(self, tensor: torch.Tensor) -> bool:
        return tensor.is_cuda and self._segment_of(tensor) is not None

    def _exchange(self, seg: _Segment) -> None:
        """Collective: map every peer's copy of this segment and bind the segment to a multicast object."""
        if seg.peer_ptrs is not None:
            return
        gathered = self.ctx.all_gather_object((seg.index, seg.nbytes))
        for rank, (index, nbytes) in enumerate(gathered):
            if index != seg.index or nbytes != seg.nbytes:
                raise RuntimeError(f"symmetric allocation order diverged between ranks (segment {seg.index}/{seg.nbytes} "
                                   f"vs rank {rank}: {index}/{nbytes})")
        peer_ptrs = [0] * self.world
        peer_ptrs[self.rank] = seg.ptr
        if self.mesh is not None:
            for peer, fd in self.mesh.all_to_all_fd(seg.fd).items():
                handle, ptr = ctypes.c_ulonglong(), ctypes.c_void_p()
                _lib.check(self.lib.fl4h_symm_map_peer(self.device_index, fd, ctypes.c_size_t(seg.nbytes),
                                                       ctypes.byref(handle), ctypes.byref(ptr)), "fl4h_symm_map_peer")
                peer_ptrs[peer] = int(ptr.value)
                seg.peer_handles.append(int(handle.value))
        os.close(seg.fd)
        seg.fd = -1
        if self.has_multicast and self.mesh is not None:
            mc = ctypes.c_ulonglong()
            fd = ctypes.c_int(-1)
            if self.rank == 0:
                _lib.check(self.lib.fl4h_symm_mc_create(self.world, ctypes.c_size_t(seg.nbytes), ctypes.byref(mc),
                                                        ctypes.byref(fd)), "fl4h_symm_mc_create")
                self.mesh.broadcast_fd(fd.value, src=0)
                os.close(fd.value)
                join_fd = -1
            else:
                join_fd = self.mesh.broadcast_fd(None, src=0)
            _lib.check(self.lib.fl4h_symm_mc_join(self.device_index, join_fd, ctypes.byref(mc)), "fl4h_symm_mc_join")
            self.ctx.barrier()  # every device is added before anybody binds
            mc_ptr = ctypes.c_void_p()
            _lib.check(self.lib.fl4h_symm_mc_bind(self.device_index, mc, ctypes.c_ulonglong(seg.handle),
                                                  ctypes.c_size_t(seg.nbytes), ctypes.byref(mc_ptr)), "fl4h_symm_mc_bind")
            seg.mc_ptr, seg.mc_handle = int(mc_ptr.value), int(mc.value)
            self.ctx.barrier()  # every rank bound its memory: multimem accesses reach all copies from here on
        seg.peer_ptrs = peer_ptrs

    def result_buffer(self, numel: int, kind: str = "result") -> torch.Tensor:
        """One cached symmetric fp32 buffer per (kind, numel): ``result`` (landing zone) or ``stage`` (pre-scaled)."""
        entry = self._result_by_key.get((kind, numel))
        if entry is None:
            entry = self._carve(numel * 4, torch.uint8)[: numel * 4].view(torch.float32)  # ONE tensor object per key
            self._result_by_key[(kind, numel)] = entry
        return entry

    # -- kernels ----------------------------------------------------------------------------------------------
    def _peer_args(self, contrib: torch.Tensor, result: torch.Tensor, coef_by_rank: list[float]) -> _PeerArgs:
        c_seg, r_seg, f_seg = self._segment_of(contrib), self._segment_of(result), self._segment_of(self._flags)
        assert c_seg is not None and r_seg is not None and f_seg is not None, "buffers must live in symmetric memory"
        self._exchange(c_seg)
        self._exchange(r_seg)
        assert c_seg.peer_ptrs is not None and r_seg.peer_ptrs is not None and f_seg.peer_ptrs is not None
        c_off, r_off, f_off = contrib.data_ptr() - c_seg.ptr, result.data_ptr() - r_seg.ptr, self._flags.data_ptr() - f_seg.ptr
        args = _PeerArgs()
        for rank in range(self.world):
            args.contrib[rank] = c_seg.peer_ptrs[rank] + c_off
            args.result[rank] = r_seg.peer_ptrs[rank] + r_off
            args.flags[rank] = f_seg.peer_ptrs[rank] + f_off
            args.coef[rank] = float(coef_by_rank[rank])
        args.rank, args.world = self.rank, self.world
        args.mc_contrib = c_seg.mc_ptr + c_off if c_seg.mc_ptr else 0
        args.mc_result = r_seg.mc_ptr + r_off if r_seg.mc_ptr else 0
        args.uniform_coef = 1.0
        return args

    def aggregate(
        self, local: torch.Tensor, coef_by_rank: list[float], out: torch.Tensor | None = None,
        epilogue: dict[str, Any] | None = None, int_local: torch.Tensor | None = None, int_out: torch.Tensor | None = None,
    ) -> torch.Tensor:
        """``result = epilogue(sum_r coef[r] * flat_r)`` on every rank, one kernel (see module docstring).

        ``int_local`` (symmetric int64 buffer, e.g. the arena's BatchNorm counters) is reduced to ``int_out`` (a plain
        local tensor, distinct from the input) by the same launch: ``int_out = trunc(sum_r coef[r] * int_r)``."""
        numel = local.numel()
        assert numel % 4 == 0 and local.dtype == torch.float32
        result = out if (out is not None and self.owns(out)) else self.result_buffer(numel)
        args = self._peer_args(local, result, coef_by_rank)
        if self.has_multicast:
            args.use_nvls = 1
            coefs = [float(c) for c in coef_by_rank[: self.world]]
            if all(c == coefs[0] for c in coefs):
                args.uniform_coef = coefs[0]  # reduce the arenas in place, scale once after the in-switch sum
            else:
                stage = self.result_buffer(numel, "stage")
                s_seg = self._segment_of(stage)
                assert s_seg is not None
                self._exchange(s_seg)
                args.stage = stage.data_ptr()
                args.mc_contrib = s_seg.mc_ptr + (stage.data_ptr() - s_seg.ptr)
        if int_local is not None and int_local.numel() > 0:
            i_seg = self._segment_of(int_local)
            assert i_seg is not None and int_out is not None and int_out.data_ptr() != int_local.data_ptr()
            assert int_local.dtype == torch.int64 and int_out.dtype == torch.int64
            self._exchange(i_seg)
            assert i_seg.peer_ptrs is not None
            i_off = int_local.data_ptr() - i_seg.ptr
            for rank in range(self.world):
                args.ibuf[rank] = i_seg.peer_ptrs[rank] + i_off
            args.ibuf_out = int_out.data_ptr()
            args.n_int = int_local.numel()
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
        args.use_nvls = 1 if self.has_multicast else 0
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
        if self.mesh is not None:
            self.mesh.close()
        # Mappings and physical memory are released with the process: tensors handed to arenas may still be
        # referenced, and unbinding a multicast object while a peer still issues multimem traffic is undefined.
        for seg in self.segments:
            if seg.fd >= 0:
                try:
                    os.close(seg.fd)
                except OSError:
                    log(WARNING, f"[fused] could not close exported fd of segment {seg.index}")
                seg.fd = -1
