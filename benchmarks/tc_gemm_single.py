"""A few launches of the tcgen05 Linear kernel on one problem size (target of the ncu captures).
    python benchmarks/tc_gemm_single.py [M N K [act]]     # default 4096 4096 4096 relu"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from fl4health_b200.ops.tc_gemm import linear_bias_act
m, n, k = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (4096, 4096, 4096)
act = sys.argv[4] if len(sys.argv) > 4 else "relu"
x = torch.randn(m, k, device="cuda").bfloat16()
w = torch.randn(n, k, device="cuda").bfloat16()
b = torch.randn(n, device="cuda")
for variant in ("4", "4", "3", "2", "1", "0"):
    os.environ["FL4H_TC_VARIANT"] = variant
    linear_bias_act(x, w, b, act)
torch.cuda.synchronize()
