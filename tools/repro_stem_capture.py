import ctypes, os, sys, torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from fl4health_b200.models import resnet18_cifar
from fl4health_b200.engine import streams
rt = ctypes.CDLL("libcudart.so.12")
def status(tag):
    st = ctypes.c_int(-1)
    rt.cudaStreamIsCapturing(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), ctypes.byref(st))
    print("   ", tag, "status", st.value, flush=True)
def run(label):
    model = resnet18_cifar().cuda().to(memory_format=torch.channels_last)
    x = torch.randn(32, 3, 32, 32, device="cuda").contiguous(memory_format=torch.channels_last)
    t = torch.randint(0, 10, (32,), device="cuda")
    for _ in range(3):
        for p in model.parameters(): p.grad = None
        F.cross_entropy(model(x), t).backward()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            for p in model.parameters(): p.grad = None
            status("begin")
            h = model.conv1(x, stats=None) if os.environ.get("STEP") == "1" else None
            out = model(x); status("after forward")
            loss = F.cross_entropy(out, t); status("after loss")
            loss.backward(); status("after backward")
        g.replay(); torch.cuda.synchronize(); print(label, "capture ok")
    except Exception as e:
        print(label, "FAILED", type(e).__name__, str(e)[:120].replace("\n", " "))
        torch.cuda.synchronize()
run(f"stem={os.environ.get('FL4H_STEM','1')} overlap={os.environ.get('FL4H_OVERLAP_WGRAD','1')}")
