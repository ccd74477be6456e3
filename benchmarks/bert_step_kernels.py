"""Per-kernel GPU time of one captured BERT-base training step (CUPTI via torch.profiler), tcgen05 LinearAct vs
stock nn.Linear (FL4H_TC_DISABLE=1): where the two paths differ."""
import collections, json, os, sys, tempfile
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))
os.environ.setdefault("FL4H_LOG_LEVEL", "WARNING")
if "--stock-linear" in sys.argv:
    os.environ["FL4H_TC_DISABLE"] = "1"
import torch
from torch import nn
from torch.profiler import ProfilerActivity, profile
import bert_fedopt_round as b
from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.engine.options import EngineOptions
from fl4health_b200.metrics import Accuracy
from fl4health_b200.models.bert import BertConfig, BertForSequenceClassification

dev = torch.device("cuda:0")
cfg = BertConfig()


class C(BasicClient):
    def get_model(self, c): return BertForSequenceClassification(cfg, 4)
    def get_data_loaders(self, c): return b.DeviceDictLoader(256, 128, cfg.vocab_size, 32, True, self.device, 1), b.DeviceDictLoader(64, 128, cfg.vocab_size, 32, False, self.device, 2)
    def get_criterion(self, c): return nn.CrossEntropyLoss()
    def get_optimizer(self, c): return torch.optim.AdamW(self.model.parameters(), lr=5e-5, weight_decay=0.01)


client = C(Path("."), [Accuracy()], dev, engine_options=EngineOptions(cuda_graphs=True, amp_dtype=torch.bfloat16, master_weights=True))
client.setup_client({"current_server_round": 1, "local_steps": 4, "batch_size": 32})
x, y = next(iter(client.train_loader))
x, y = client._prepare_batch(x, y)
client.model.train()
for _ in range(8):
    client._run_train_unit(x, y)
torch.cuda.synchronize()
N = 4
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        client._run_train_unit(x, y)
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "bert_trace.json")
prof.export_chrome_trace(path)
events = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
events.sort(key=lambda e: e["ts"])
per = collections.defaultdict(lambda: [0, 0.0])
for e in events:
    per[e["name"][:90]][0] += 1
    per[e["name"][:90]][1] += e["dur"]
total = sum(v[1] for v in per.values())
span = events[-1]["ts"] + events[-1]["dur"] - events[0]["ts"]
print(f"steps={N} kernels/step={len(events)/N:.0f} sum_kernel_ms/step={total/N/1e3:.3f} span_ms/step={span/N/1e3:.3f}")
for name, (count, dur) in sorted(per.items(), key=lambda kv: -kv[1][1])[:32]:
    print(f"{dur/N:9.1f} us {count/N:6.1f} x {dur/count:8.2f}  {name}")
