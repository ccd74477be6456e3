"""Phase timeline of CTA 0 of the persistent tcgen05 Linear kernel (SM-clock timestamps written by the kernel itself):
where a small-K GEMM spends its time.   python benchmarks/tc_gemm_phases.py [M N K [variant]]"""
import ctypes, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from fl4health_b200.ops import _lib
from fl4health_b200.ops.tc_gemm import linear_bias_act
m, n, k = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (4096, 2304, 768)
os.environ["FL4H_TC_VARIANT"] = sys.argv[4] if len(sys.argv) > 4 else "3"
x = torch.randn(m, k, device="cuda").bfloat16()
w = torch.randn(n, k, device="cuda").bfloat16()
b = torch.randn(n, device="cuda")
for _ in range(20):
    linear_bias_act(x, w, b, "relu")
torch.cuda.synchronize()
start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
start.record(); linear_bias_act(x, w, b, "relu"); end.record(); torch.cuda.synchronize()
buf = (ctypes.c_longlong * 48)()
assert _lib.load(True).fl4h_tc_debug_read(buf) == 0
t = list(buf)
us = lambda a, c: (t[a] - t[c]) / 1965.0  # noqa: E731  (SM clock 1965 MHz)
print(f"shape {m}x{n}x{k} variant {os.environ['FL4H_TC_VARIANT']}: event time {start.elapsed_time(end)*1e3:.1f} us (includes host launch)")
print(f"  setup (barriers + TMEM alloc + sync)      {us(1, 0):6.2f} us")
print(f"  first operands landed (TMA latency)       {us(2, 1):6.2f} us after setup")
tiles = sum(1 for i in range(8) if t[3 + i] > t[0])
prev = 2
for i in range(tiles):
    print(f"  tile {i}: MMA issue done at {us(3 + i, 0):6.2f} us (+{us(3 + i, prev):5.2f}) | accumulator visible {us(16 + 2*i, 0):6.2f} | epilogue done {us(17 + 2*i, 0):6.2f} (epilogue {us(17 + 2*i, 16 + 2*i):5.2f} us)")
    prev = 3 + i
print(f"  tile 0, epilogue warp 0, first 32x32 chunk: tcgen05.ld+wait {us(43, 42):5.2f} us, bias+activation {us(44, 43):5.2f}, pack+STS+syncwarp {us(45, 44):5.2f}, LDS {us(46, 45):5.2f}, STG+syncwarp {us(47, 46):5.2f} us")
print(f"  last TMA issued at {us(40, 0):6.2f} us; kernel exit at {us(41, 0):6.2f} us")
