"""One launch of each tcgen05 Linear variant on a 4096^3 problem (target of the ncu capture)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from fl4health_b200.ops.tc_gemm import linear_bias_act
x = torch.randn(4096, 4096, device="cuda").bfloat16()
w = torch.randn(4096, 4096, device="cuda").bfloat16()
b = torch.randn(4096, device="cuda")
for variant in ("2", "2", "1", "0"):
    os.environ["FL4H_TC_VARIANT"] = variant
    linear_bias_act(x, w, b, True)
torch.cuda.synchronize()
