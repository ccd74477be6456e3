"""Diagnostic: per-parameter gradient agreement of ResNet-18 between the fused-BN path and the stock-op fallback."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import fl4health_b200.ops.bn_act as bn_mod
from fl4health_b200.models import resnet18_cifar

torch.manual_seed(0)
dev = torch.device("cuda")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
model = resnet18_cifar().to(dev).to(memory_format=torch.channels_last)
x = torch.randn(16, 3, 32, 32, device=dev).contiguous(memory_format=torch.channels_last)
target = torch.randint(0, 10, (16,), device=dev)
state = {k: v.clone() for k, v in model.state_dict().items()}

def run(fused):
    model.load_state_dict(state)
    model.zero_grad()
    original = bn_mod.kernel_eligible
    if not fused:
        bn_mod.kernel_eligible = lambda *a, **k: False
    acts = {}
    hooks = [m.register_forward_hook(lambda mod, i, o, n=n: acts.__setitem__(n, o.detach().float().clone())) for n, m in model.named_modules() if n.endswith(("bn1", "bn2", "downsample.1"))]
    try:
        loss = torch.nn.functional.cross_entropy(model(x), target)
        loss.backward()
    finally:
        bn_mod.kernel_eligible = original
        for h in hooks: h.remove()
    torch.cuda.synchronize()
    return loss.item(), {n: p.grad.clone() for n, p in model.named_parameters()}, acts

for overlap in ("0", "1"):
    os.environ["FL4H_OVERLAP_WGRAD"] = overlap
    l0, g0, a0 = run(False)
    l1, g1, a1 = run(True)
    print(f"overlap={overlap} loss {l0:.6f} {l1:.6f}")
    for n in a0:
        err = float((a0[n] - a1[n]).abs().max() / a0[n].abs().max().clamp_min(1e-6))
        if err > 1e-4: print("  act", n, err)
    for n in g0:
        err = float((g0[n] - g1[n]).abs().max() / g0[n].abs().max().clamp_min(1e-9))
        print(f"  {n:40s} {err:.2e}  |g|max {float(g0[n].abs().max()):.3e}")
