"""Throughput of the hand-written tcgen05 Linear kernel vs cuBLAS (torch.nn.functional.linear), bf16, CUDA events."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import os
from fl4health_b200.ops.tc_gemm import linear_bias_act

def timed(fn, iters=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(iters):
        fn()
    end.record()
    torch.cuda.synchronize()
    return start.elapsed_time(end) / iters

dev = torch.device("cuda")
print(f"{'M':>6} {'N':>6} {'K':>6} | tcgen05 TFLOP/s: v0 (1 tile/CTA)  v1 (persistent 128x128)  v2 (persistent 128x256) | cuBLAS+bias+relu TFLOP/s")
for m, n, k in [(256, 512, 512), (1024, 1024, 1024), (4096, 4096, 4096), (8192, 8192, 8192), (16384, 4096, 1024),
                (4096, 2304, 768), (4096, 768, 768), (4096, 3072, 768), (4096, 768, 3072)]:  # last four: BERT-base, 32 x 128 tokens
    x = torch.randn(m, k, device=dev).bfloat16()
    w = torch.randn(n, k, device=dev).bfloat16()
    b = torch.randn(n, device=dev)
    bb = b.bfloat16()
    flops = 2.0 * m * n * k
    ours = []
    for variant in ("0", "1", "2"):
        os.environ["FL4H_TC_VARIANT"] = variant
        ours.append(flops / timed(lambda: linear_bias_act(x, w, b, True)) / 1e9)
    lib = timed(lambda: torch.relu(torch.nn.functional.linear(x, w, bb)))
    print(f"{m:6d} {n:6d} {k:6d} | {ours[0]:10.1f} {ours[1]:10.1f} {ours[2]:10.1f} | {flops / lib / 1e9:8.1f}")
