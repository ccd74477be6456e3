"""In-situ kernel timeline of the captured training step (CUPTI through torch.profiler; no cache flushing, real
inter-kernel gaps): per-kernel totals plus stream occupancy of one graph replay."""
import json, os, sys, collections, tempfile
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import torch
from torch.profiler import ProfilerActivity, profile
import profile_step as ps

device = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
from fl4health_b200.engine.options import EngineOptions
from fl4health_b200.metrics import Accuracy
BF16 = os.environ.get("KINETO_DTYPE", "bf16") == "bf16"
engine = EngineOptions(cuda_graphs=True, amp_dtype=torch.bfloat16 if BF16 else None, channels_last=True, master_weights=BF16,
                       table_grads=not BF16)
KIND = os.environ.get("KINETO_CLIENT", "basic")  # basic | fedprox : which client algorithm's step is profiled
if KIND == "fedprox":
    from fl4health_b200.clients.fed_prox_client import FedProxClient

    client_cls = type("ProfFedProx", (ps.Client, FedProxClient), {"get_model": ps.Client.get_model})
else:
    client_cls = ps.Client
client = client_cls(Path("."), [Accuracy()], device, client_name="prof", engine_options=engine)
cfg = {"current_server_round": 1, "local_steps": 8, "batch_size": ps.BS}
client.setup_client(cfg)
if KIND == "fedprox":
    client.drift_penalty_weight = 0.1
    client.drift_penalty_tensors = client.snapshot_drift_anchor()
client.model.train()
x, y = next(iter(client.train_loader))
x, y = client._prepare_batch(x, y)
for _ in range(8):
    client._run_train_unit(x, y)
torch.cuda.synchronize()
if os.environ.get("KINETO_GC") == "1":  # what keeps autograd history alive between steps (it should be nothing)
    import gc

    gc.collect()
    alive = [o for o in gc.get_objects() if isinstance(o, torch.Tensor) and o.grad_fn is not None]
    print(f"tensors with autograd history alive after a step: {len(alive)}")
    for t in alive[:12]:
        holders = [type(r).__name__ + (":" + ",".join(k for k, v in r.items() if v is t)[:60] if isinstance(r, dict) else "") for r in gc.get_referrers(t)][:4]
        print("  ", tuple(t.shape), t.dtype, type(t.grad_fn).__name__, holders)
N = 5
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(N):
        client._run_train_unit(x, y)
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "trace.json")
prof.export_chrome_trace(path)
events = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
events.sort(key=lambda e: e["ts"])
per = collections.defaultdict(lambda: [0, 0.0])
for e in events:
    per[e["name"][:70]][0] += 1
    per[e["name"][:70]][1] += e["dur"]
total = sum(v[1] for v in per.values())
span = (events[-1]["ts"] + events[-1]["dur"] - events[0]["ts"])
print(f"replays={N} kernels/replay={len(events)/N:.0f} sum_kernel_us/replay={total/N:.1f} span_us/replay={span/N:.1f}")
streams = collections.defaultdict(float)
for e in events:
    streams[e["args"].get("stream")] += e["dur"]
print("busy us per replay by stream:", {k: round(v / N, 1) for k, v in streams.items()})
for name, (count, dur) in sorted(per.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f"{dur/N:8.1f} us {count/N:5.1f} x {dur/count:6.2f}  {name}")
# critical-path view: gaps on the main stream
main = max(streams, key=streams.get)
ev = [e for e in events if e["args"].get("stream") == main]
gaps = [b["ts"] - (a["ts"] + a["dur"]) for a, b in zip(ev, ev[1:])]
gaps = [g for g in gaps if g < 200]
print(f"main stream: {len(ev)/N:.0f} kernels/replay, mean gap {sum(gaps)/len(gaps):.2f} us, total gap/replay {sum(gaps)/N:.1f} us")
