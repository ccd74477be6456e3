"""Diagnostic: SCAFFOLD example under SPMD with per-round norms of what the strategy aggregates and produces."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from fl4health_b200.strategies import scaffold as S

orig = S.Scaffold.aggregate


def aggregate(self, params):
    rank = os.environ.get("RANK")
    for p in params:
        w, c = self.parameter_packer.unpack_parameters(p)
        local = getattr(p, "rank", None) == int(rank)
        if local:
            print(f"DBG r{rank} local payload: w_norm={float(torch.cat([t.float().flatten() for t in w if t.is_floating_point()]).norm()):.5f} "
                  f"dc_norm={float(torch.cat([t.float().flatten() for t in c]).norm()):.5f} specs={w.spec.flat_numel},{c.spec.flat_numel}", flush=True)
    out = orig(self, params)
    w, c = self.parameter_packer.unpack_parameters(out)
    print(f"DBG r{rank} aggregated: w_norm={float(torch.cat([t.float().flatten() for t in w if t.is_floating_point()]).norm()):.5f} "
          f"dc_norm={float(torch.cat([t.float().flatten() for t in c]).norm()):.5f} flat={getattr(w, 'flat', None) is not None},{getattr(c, 'flat', None) is not None}", flush=True)
    return out


S.Scaffold.aggregate = aggregate

from fl4health_b200.parallel import spmd as P
import torch.distributed as dist

orig_ws = P.SpmdContext.weighted_sum_flat


def checked(self, local, coef_by_rank, numel, out=None, epilogue=None, int_local=None):
    expect = None
    if local is not None and self.world_size > 1 and not epilogue:
        expect = local[:numel].detach().clone() * coef_by_rank[self.rank]
        dist.all_reduce(expect)
    got = orig_ws(self, local, coef_by_rank, numel, out=out, epilogue=epilogue, int_local=int_local)
    if expect is not None:
        torch.cuda.synchronize()
        print(f"DBG r{self.rank} weighted_sum_flat numel={numel} fused={self.fused is not None and self.fused.owns(local)} "
              f"max|fused-nccl|={float((got[:numel] - expect).abs().max()):.3e} |expect|={float(expect.norm()):.4f} |got|={float(got[:numel].norm()):.4f} "
              f"local_ptr_off={local.data_ptr() % (1 << 28)}", flush=True)
    return got


P.SpmdContext.weighted_sum_flat = checked
from examples.run import main

main(["scaffold_example", "--spmd", "--rounds", "3"])
