"""Throughput of the hand-written tcgen05 Linear kernel vs cuBLAS (torch.nn.functional.linear), bf16, CUDA events."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import os
from fl4health_b200.ops.tc_gemm import linear_bias_act

def timed(fn, iters=20, repeats=3):
    """Best of `repeats` blocks of `iters` back-to-back calls (CUDA events).  The first version of this script measured
    each configuration once, cold: the SM clock was still ramping during the first configurations of every shape and
    the ORDER of the columns changed the ranking.  `_spin` keeps the clocks up between configurations."""
    best = float("inf")
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    # the calls are captured in a CUDA graph: small problems are otherwise HOST-bound (ctypes call + two tensor-map
    # encodes + autograd.Function cost ~23 us per call -- every small shape used to report the same "23 us kernel")
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for _ in range(iters):
                fn()
    for _ in range(repeats):
        _spin()
        graph.replay()
        torch.cuda.synchronize()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        graph.replay()
        end.record()
        torch.cuda.synchronize()
        best = min(best, start.elapsed_time(end) / iters)
    return best


_SPIN = None


def _spin():
    global _SPIN
    if _SPIN is None:
        _SPIN = (torch.randn(4096, 4096, device="cuda").bfloat16(), torch.randn(4096, 4096, device="cuda").bfloat16())
    for _ in range(40):  # ~4 ms of dense tensor-core work
        _SPIN[0] @ _SPIN[1]

dev = torch.device("cuda")
print(f"{'M':>6} {'N':>6} {'K':>6} | tcgen05 TFLOP/s: v0 (1 tile/CTA)  v1 (persistent 128x128)  v2 (persistent 128x256)  v3 (v2, 8 epilogue warps)  v4 (CTA pair) | GELU epilogue | library")
for m, n, k in [(256, 512, 512), (1024, 1024, 1024), (4096, 4096, 4096), (8192, 8192, 8192), (16384, 4096, 1024),
                (4096, 2304, 768), (4096, 768, 768), (4096, 3072, 768), (4096, 768, 3072)]:  # last four: BERT-base, 32 x 128 tokens
    x = torch.randn(m, k, device=dev).bfloat16()
    w = torch.randn(n, k, device=dev).bfloat16()
    b = torch.randn(n, device=dev)
    bb = b.bfloat16()
    flops = 2.0 * m * n * k
    ours = []
    for variant in ("0", "1", "2", "3", "4"):
        os.environ["FL4H_TC_VARIANT"] = variant
        ours.append(flops / timed(lambda: linear_bias_act(x, w, b, True)) / 1e9)
    lib = timed(lambda: torch.relu(torch.nn.functional.linear(x, w, bb)))
    os.environ["FL4H_TC_VARIANT"] = "2"
    gelu2 = flops / timed(lambda: linear_bias_act(x, w, b, "gelu")) / 1e9
    os.environ["FL4H_TC_VARIANT"] = "3"
    gelu3 = flops / timed(lambda: linear_bias_act(x, w, b, "gelu")) / 1e9
    os.environ["FL4H_TC_VARIANT"] = "4"
    gelu4 = flops / timed(lambda: linear_bias_act(x, w, b, "gelu")) / 1e9
    lib_gemm = timed(lambda: torch.nn.functional.linear(x, w, bb))
    lib_gelu = timed(lambda: torch.nn.functional.gelu(torch.nn.functional.linear(x, w, bb)))
    print(f"{m:6d} {n:6d} {k:6d} | {ours[0]:8.1f} {ours[1]:8.1f} {ours[2]:8.1f} {ours[3]:8.1f} {ours[4]:8.1f} | gelu v2 {gelu2:7.1f} v3 {gelu3:7.1f} v4 {gelu4:7.1f} | "
          f"lib gemm+bias {flops / lib_gemm / 1e9:7.1f}  +relu {flops / lib / 1e9:7.1f}  +gelu {flops / lib_gelu / 1e9:7.1f}")
