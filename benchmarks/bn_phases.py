"""Phase timeline of CTA 0 of the fused BatchNorm(+ReLU) forward kernel (SM-clock timestamps written by the kernel):
which of load / block-reduce+atomics / grid barrier / totals / store owns its ~10 us.
    python benchmarks/bn_phases.py [N C H W]"""
import ctypes, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from fl4health_b200.ops import _lib
from fl4health_b200.ops.bn_act import batch_norm_act
n, c, h, w = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (32, 64, 32, 32)
x = torch.randn(n, c, h, w, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
gamma, beta = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
rm, rv, nbt = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda"), torch.zeros((), dtype=torch.int64, device="cuda")
flush = torch.empty(64 << 20, device="cuda")
for cold in (False, True):
    for _ in range(10):
        batch_norm_act(x, gamma, beta, rm, rv, nbt, True, 0.1, 1e-5, None, True)
    if cold:
        flush.fill_(1.0)  # evict x from L2
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record(); batch_norm_act(x, gamma, beta, rm, rv, nbt, True, 0.1, 1e-5, None, True); end.record()
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 16)()
    assert _lib.load(True).fl4h_bn_debug_read(buf) == 0
    t = list(buf)
    us = lambda a, b: (t[a] - t[b]) / 1965.0  # noqa: E731
    print(f"{n}x{c}x{h}x{w} bf16 NHWC ({x.numel()*2/1e6:.1f} MB) {'cold' if cold else 'L2-warm'}: event {start.elapsed_time(end)*1e3:.1f} us (with launch); CTA 0 total {us(6, 0):.2f} us")
    print(f"   stage shift values {us(1, 0):.2f} | load tile + sums {us(2, 1):.2f} | block reduce + RED atomics {us(3, 2):.2f} | grid barrier {us(4, 3):.2f} | totals -> scale/shift {us(5, 4):.2f} | normalise + store {us(6, 5):.2f}")

# backward kernel (ReLU mask, no residual): forward once with autograd, time the backward launch
y = batch_norm_act(x, gamma.requires_grad_(True), beta.requires_grad_(True), rm, rv, nbt, True, 0.1, 1e-5, None, True)
g = torch.randn_like(y)
for _ in range(5):
    y.backward(g, retain_graph=True)
torch.cuda.synchronize()
y.backward(g, retain_graph=True)
torch.cuda.synchronize()
assert _lib.load(True).fl4h_bn_debug_read(buf) == 0
t = list(buf)
print(f"backward, L2-warm: CTA 0 total {us(13, 8):.2f} us")
print(f"   load tile + sums {us(9, 8):.2f} | block reduce + RED atomics {us(10, 9):.2f} | grid barrier {us(11, 10):.2f} | totals staged {us(12, 11):.2f} | dx from registers + store {us(13, 12):.2f}")
