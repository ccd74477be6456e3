"""Timing decomposition of the fused collectives (torchrun, >= 2 GPUs): fixed overhead (tiny payload), pure transfer
(broadcast without unpack targets), full kernels; one JSON line per configuration.  Knobs come from the environment
(FL4H_NVLS, FL4H_NVLS_UNROLL, FL4H_COLL_GRID) so a shell loop sweeps them."""

import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ.setdefault("FL4H_LOG_LEVEL", "WARNING")

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from fl4health_b200.parallel.spmd import SpmdContext  # noqa: E402


def main() -> None:
    ctx = SpmdContext()
    assert ctx.enable_fused_collectives()
    fused, rank, world, dev = ctx.fused, ctx.rank, ctx.world_size, ctx.device
    numel = int(os.environ.get("TUNE_NUMEL", 11173888))
    local = fused.allocator(numel, torch.float32, dev)
    local.normal_()
    tiny = fused.allocator(4096, torch.float32, dev)
    uni = [1.0 / world] * world
    skew = [(r + 1) / sum(range(1, world + 1)) for r in range(world)]
    w = torch.empty(numel, device=dev)
    scratch = torch.empty(numel, device=dev)

    def timed(fn, iters=30):  # noqa: ANN001, ANN202
        for _ in range(5):
            fn()
        ctx.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        return round(ctx.all_reduce_max(s.elapsed_time(e) / iters) * 1e3, 1)  # microseconds

    out = {
        "world": world, "nvls": bool(fused.has_multicast), "unroll": os.environ.get("FL4H_NVLS_UNROLL", "4"),
        "grid": os.environ.get("FL4H_COLL_GRID", "sms"), "payload_MB": numel * 4 / 1e6,
        "us_agg_tiny": timed(lambda: fused.aggregate(tiny, uni)),
        "us_agg_uniform": timed(lambda: fused.aggregate(local, uni)),
        "us_agg_weighted": timed(lambda: fused.aggregate(local, skew)),
        "us_bcast_tiny": timed(lambda: fused.broadcast(tiny, 0)),
        "us_bcast_transfer_only": timed(lambda: fused.broadcast(local, 0)),
        "us_bcast_plus_w": timed(lambda: fused.broadcast(local, 0, w=w)),
        "us_nccl_allreduce": timed(lambda: dist.all_reduce(scratch)),
        "us_nccl_bcast": timed(lambda: dist.broadcast(scratch, src=0)),
    }
    if fused.has_multicast:  # kernel-written phase timeline of the last launch (CTA 0, ns since entry)
        import ctypes

        for label, fn in (("tiny", lambda: fused.aggregate(tiny, uni)), ("uniform", lambda: fused.aggregate(local, uni)),
                          ("weighted", lambda: fused.aggregate(local, skew))):
            fn()
            torch.cuda.synchronize()
            buf = (ctypes.c_ulonglong * 8)()
            fused.lib.fl4h_coll_debug_read(buf)
            out[f"phases_us_{label}"] = [round((buf[i] - buf[0]) / 1e3, 1) for i in range(1, 7)]
    if rank == 0:
        print("TUNE " + json.dumps(out))
    ctx.barrier()
    ctx.shutdown()


if __name__ == "__main__":
    main()
