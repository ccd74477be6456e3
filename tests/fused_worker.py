"""Multi-GPU worker: correctness + bandwidth of the fused peer-memory collectives (torch.distributed.run, nccl)."""

import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("FL4H_LOG_LEVEL", "WARNING")

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from fl4health_b200.ops import flat as F  # noqa: E402
from fl4health_b200.parallel.spmd import SpmdContext  # noqa: E402


def main() -> None:
    out_path = sys.argv[1]
    numel = int(sys.argv[2]) if len(sys.argv) > 2 else 11_200_000 // 32 * 32
    ctx = SpmdContext()
    assert ctx.enable_fused_collectives(), "fused collectives could not be enabled"
    fused = ctx.fused
    rank, world, dev = ctx.rank, ctx.world_size, ctx.device
    report: dict = {"world": world, "numel": numel}

    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    local = fused.allocator(numel, torch.float32, dev)
    local.copy_(torch.randn(numel, generator=gen, device=dev))
    coefs = [(r + 1) / sum(range(1, world + 1)) for r in range(world)]

    # reference: gather everything with NCCL and reduce in the same fixed order
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    ref = torch.zeros_like(local)
    F.weighted_sum(ref, gathered, coefs)

    out = fused.aggregate(local, coefs)
    torch.cuda.synchronize()
    report["agg_max_abs_err"] = float((out - ref).abs().max().item())
    report["agg_bit_exact"] = bool(torch.equal(out, ref))

    # FedAdam epilogue: compare with the single-GPU epilogue kernel on the gathered data
    cur = torch.randn(numel, generator=torch.Generator(device=dev).manual_seed(7), device=dev)
    m_ref, v_ref = torch.zeros(numel, device=dev), torch.zeros(numel, device=dev)
    ref_adam = torch.empty(numel, device=dev)
    kw = dict(mode=F.EPI_FEDADAM, eta=0.1, beta1=0.9, beta2=0.99, tau=1e-3)
    F.weighted_sum(ref_adam, gathered, coefs, current=cur, m=m_ref, v=v_ref, **kw)
    m, v = torch.zeros(numel, device=dev), torch.zeros(numel, device=dev)
    out_adam = fused.aggregate(local, coefs, epilogue=dict(current=cur, m=m, v=v, **kw)).clone()
    torch.cuda.synchronize()
    report["adam_max_abs_err"] = float((out_adam - ref_adam).abs().max().item())

    # broadcast from the last rank + fused unpack (w, anchor, bf16 shadow, scaffold correction)
    root = world - 1
    w, anchor = torch.empty(numel, device=dev), torch.empty(numel, device=dev)
    shadow = torch.empty(numel, device=dev, dtype=torch.bfloat16)
    c_server, c_local, cv = torch.randn(numel, device=dev), torch.randn(numel, device=dev), torch.empty(numel, device=dev)
    fused.broadcast(local, root, w=w, anchor=anchor, shadow=shadow, c_server=c_server, c_local=c_local, cv_out=cv)
    torch.cuda.synchronize()
    expected = gathered[root]
    report["bcast_ok"] = bool(torch.equal(w, expected) and torch.equal(anchor, expected)
                              and torch.equal(shadow, expected.to(torch.bfloat16)) and torch.allclose(cv, c_server - c_local))

    # timing (device events, max over ranks): fused vs NCCL all-reduce / broadcast of the same payload
    def timed(fn, iters=20):  # noqa: ANN001, ANN202
        for _ in range(3):
            fn()
        ctx.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        return ctx.all_reduce_max(s.elapsed_time(e) / iters)

    scratch = torch.empty_like(local)

    def nccl_allreduce() -> None:
        F.weighted_sum(scratch, [local], [coefs[rank]])
        dist.all_reduce(scratch)

    def nccl_bcast() -> None:
        dist.broadcast(scratch, src=root)
        F.bcast_unpack(scratch, w, anchor, shadow, c_server, c_local, cv)

    report["ms_fused_agg"] = timed(lambda: fused.aggregate(local, coefs))
    report["ms_nccl_agg"] = timed(nccl_allreduce)
    report["ms_fused_bcast"] = timed(lambda: fused.broadcast(local, root, w=w, anchor=anchor, shadow=shadow,
                                                             c_server=c_server, c_local=c_local, cv_out=cv))
    report["ms_nccl_bcast"] = timed(nccl_bcast)
    bytes_link = numel * 4 * (world - 1) / world
    report["fused_agg_GBps_per_dir"] = bytes_link / report["ms_fused_agg"] / 1e6
    report["fused_bcast_GBps_per_dir"] = bytes_link / report["ms_fused_bcast"] / 1e6
    if rank == 0:
        Path(out_path).write_text(json.dumps(report, indent=1))
        print(json.dumps(report))
    ctx.barrier()
    ctx.shutdown()


if __name__ == "__main__":
    main()
