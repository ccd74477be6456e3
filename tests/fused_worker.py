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
    report: dict = {"world": world, "numel": numel, "nvls": bool(fused.has_multicast)}

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
    report["agg_bit_exact"] = bool(torch.equal(out, ref))  # guaranteed by the fixed-order P2P kernel only

    # uniform weights (NVLS: arenas reduced in place, no staging pass) + integer buffers riding along
    uni = [1.0 / world] * world
    ref_uni = torch.zeros_like(local)
    F.weighted_sum(ref_uni, gathered, uni)
    ints = fused.allocator(24, torch.int64, dev)
    ints.copy_(torch.arange(24, device=dev) * (rank + 1) + 7)
    int_out = torch.zeros(24, dtype=torch.int64, device=dev)
    out_uni = fused.aggregate(local, uni, int_local=ints, int_out=int_out).clone()
    torch.cuda.synchronize()
    report["agg_uniform_max_abs_err"] = float((out_uni - ref_uni).abs().max().item())
    want_int = sum((torch.arange(24, dtype=torch.float64) * (r + 1) + 7) * uni[r] for r in range(world)).to(torch.int64)
    report["int_ok"] = bool(torch.equal(int_out.cpu(), want_int))

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
    report["ms_fused_agg_uniform"] = timed(lambda: fused.aggregate(local, uni))
    report["ms_nccl_agg"] = timed(nccl_allreduce)
    report["ms_fused_bcast"] = timed(lambda: fused.broadcast(local, root, w=w, anchor=anchor, shadow=shadow,
                                                             c_server=c_server, c_local=c_local, cv_out=cv))
    report["ms_nccl_bcast"] = timed(nccl_bcast)
    # roofline: NVLink 5 = 900 GB/s per direction per GPU.
    #   aggregate: NVLS moves payload*(1 + 1/K) per direction per GPU (ld_reduce pulls every rank's slice through the
    #   switch, multimem.st pushes 1/K out and receives the whole result); P2P moves 2*payload*(K-1)/K.
    #   broadcast: the root emits the payload once (multicast) / (K-1)/K of it per peer pair (P2P scatter+all-gather).
    payload = numel * 4
    agg_bytes = payload * (1 + 1 / world) if fused.has_multicast else 2 * payload * (world - 1) / world
    bcast_bytes = payload if fused.has_multicast else payload * (world - 1) / world
    report["payload_MB"] = payload / 1e6
    for key, nbytes in (("ms_fused_agg", agg_bytes), ("ms_fused_agg_uniform", agg_bytes), ("ms_fused_bcast", bcast_bytes)):
        gbps = nbytes / report[key] / 1e6
        report[key.replace("ms_", "") + "_GBps_per_dir"] = gbps
        report[key.replace("ms_", "") + "_frac_of_900GBps"] = gbps / 900.0
    report["agg_lower_bound_ms"] = agg_bytes / 900e9 * 1e3
    report["bcast_lower_bound_ms"] = bcast_bytes / 900e9 * 1e3
    if rank == 0:
        Path(out_path).write_text(json.dumps(report, indent=1))
        print(json.dumps(report))
    ctx.barrier()
    ctx.shutdown()


if __name__ == "__main__":
    main()
