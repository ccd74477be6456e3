"""In-situ timeline of whole FL rounds of the bench configuration (1 GPU, torch.profiler / CUPTI): how much of a round
the GPU is busy, and which host phase owns each idle gap.  Diagnostic only (profiling inflates host time)."""
import collections, json, os, sys, tempfile
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ.setdefault("FL4H_LOG_LEVEL", "WARNING")
import torch
from torch import nn
from torch.profiler import ProfilerActivity, profile
from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.engine.data import BatchedTensorLoader
from fl4health_b200.engine.options import EngineOptions
from fl4health_b200.metrics import Accuracy
from fl4health_b200.metrics.metric_aggregation import evaluate_metrics_aggregation_fn, fit_metrics_aggregation_fn
from fl4health_b200.models import resnet18_cifar
from fl4health_b200.servers.base_server import FlServer
from fl4health_b200.servers.client_manager import SimpleClientManager
from fl4health_b200.simulation import register_clients
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg
from fl4health_b200.utils.dataset import TensorDataset

dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
WARM, ROUNDS = 4, 3


class C(BasicClient):
    def get_model(self, config): return resnet18_cifar()
    def get_data_loaders(self, config):
        ds = TensorDataset(torch.randn(4096, 3, 32, 32), torch.randint(0, 10, (4096,)))
        vs = TensorDataset(torch.randn(128, 3, 32, 32), torch.randint(0, 10, (128,)))
        return (BatchedTensorLoader(ds, 32, shuffle=True, drop_last=True, placement="device", device=self.device),
                BatchedTensorLoader(vs, 32, placement="device", device=self.device))
    def get_criterion(self, config): return nn.CrossEntropyLoss()
    def get_optimizer(self, config): return torch.optim.SGD(self.model.parameters(), lr=0.01, momentum=0.9)


def cfg(r): return {"current_server_round": r, "local_steps": 8, "batch_size": 32}


client = C(Path("."), [Accuracy()], dev, client_name="k", engine_options=EngineOptions(cuda_graphs=True, amp_dtype=torch.bfloat16, channels_last=True, master_weights=True))
strategy = BasicFedAvg(min_fit_clients=1, min_evaluate_clients=1, min_available_clients=1, on_fit_config_fn=cfg, on_evaluate_config_fn=cfg,
                       fit_metrics_aggregation_fn=fit_metrics_aggregation_fn, evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn)
server = FlServer(SimpleClientManager(), {"n_server_rounds": WARM + ROUNDS}, strategy, on_init_parameters_config_fn=cfg, accept_failures=False)
register_clients(server, [client])
prof = profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU])


def hook(server_round: int) -> None:
    if server_round == WARM:
        torch.cuda.synchronize()
        prof.__enter__()
    elif server_round == WARM + ROUNDS:
        torch.cuda.synchronize()
        prof.__exit__(None, None, None)


server.round_end_hooks = [hook]
server.fit(num_rounds=WARM + ROUNDS)
path = os.path.join(tempfile.gettempdir(), "round_trace.json")
prof.export_chrome_trace(path)
trace = json.load(open(path))["traceEvents"]
gpu = sorted((e for e in trace if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")), key=lambda e: e["ts"])
notes = sorted((e for e in trace if e.get("cat") == "user_annotation" and e["name"].startswith("fl4h:")), key=lambda e: e["ts"])
# merge busy intervals over all streams
merged = []
for e in gpu:
    a, b = e["ts"], e["ts"] + e["dur"]
    if merged and a <= merged[-1][1]:
        merged[-1][1] = max(merged[-1][1], b)
    else:
        merged.append([a, b])
span = merged[-1][1] - merged[0][0]
busy = sum(b - a for a, b in merged)
print(f"rounds={ROUNDS} span_ms/round={span/ROUNDS/1e3:.3f} gpu_busy_ms/round={busy/ROUNDS/1e3:.3f} idle_ms/round={(span-busy)/ROUNDS/1e3:.3f} kernels/round={len(gpu)/ROUNDS:.0f}")


def owner(ts: float) -> str:
    best = "between phases"
    for n in notes:
        if n["ts"] <= ts <= n["ts"] + n["dur"]:
            best = n["name"]
    return best


idle = collections.defaultdict(lambda: [0, 0.0])
big = []
for (a0, b0), (a1, b1) in zip(merged, merged[1:]):
    gap = a1 - b0
    if gap <= 0:
        continue
    key = owner(b0 + gap / 2)
    idle[key][0] += 1
    idle[key][1] += gap
    if gap > 25:
        big.append((gap, key, b0 - merged[0][0]))
for key, (count, total) in sorted(idle.items(), key=lambda kv: -kv[1][1]):
    print(f"idle {total/ROUNDS:8.1f} us/round in {count/ROUNDS:6.1f} gaps  while host is in: {key}")
busy_by = collections.defaultdict(float)
for e in gpu:
    busy_by[owner(e["ts"])] += e["dur"]
for key, total in sorted(busy_by.items(), key=lambda kv: -kv[1]):
    print(f"kernel time {total/ROUNDS:8.1f} us/round launched for: {key}")
print("largest gaps (us, host phase, offset ms):")
for gap, key, off in sorted(big, reverse=True)[:25]:
    print(f"  {gap:8.1f}  {key:28s} @ {off/1e3:8.3f}")
# what the host was doing inside the largest steady-state gaps (second profiled round onwards)
host = sorted((e for e in trace if e.get("cat") in ("cpu_op", "cuda_runtime", "cuda_driver", "user_annotation") and "dur" in e), key=lambda e: e["ts"])
t0 = merged[0][0]
steady = [g for g in sorted(big, reverse=True) if g[2] > span / ROUNDS][:8]
for gap, key, off in steady:
    a, b = t0 + off, t0 + off + gap
    inside = [e for e in host if e["ts"] < b and e["ts"] + e["dur"] > a and e["dur"] > 4]
    inside.sort(key=lambda e: -e["dur"])
    print(f"--- gap {gap:.0f} us in {key} @ {off/1e3:.3f} ms: host events overlapping it")
    for e in inside[:14]:
        print(f"      {e['dur']:8.1f} us  [{e['cat']}] {e['name'][:90]}  (starts {e['ts']-a:+.0f} us)")
