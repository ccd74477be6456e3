"""Wall-clock breakdown of one FL round of the bench configuration (1 GPU, synchronising at phase boundaries).
Diagnostic only: numbers include the syncs it inserts."""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ.setdefault("FL4H_LOG_LEVEL", "WARNING")
import torch
from torch import nn
from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.common.typing import ndarrays_to_parameters, parameters_to_ndarrays, NDArrays
from fl4health_b200.engine.data import BatchedTensorLoader
from fl4health_b200.engine.options import EngineOptions
from fl4health_b200.metrics import Accuracy
from fl4health_b200.models import resnet18_cifar
from fl4health_b200.strategies.aggregate_utils import aggregate_results
from fl4health_b200.utils.dataset import TensorDataset

dev = torch.device("cuda:0")
PLACEMENT = os.environ.get("PLACEMENT", "device")  # "pinned": every batch is staged from page-locked host memory
torch.backends.cudnn.benchmark = True
class C(BasicClient):
    def get_model(self, config): return resnet18_cifar()
    def get_data_loaders(self, config):
        ds = TensorDataset(torch.randn(4096, 3, 32, 32), torch.randint(0, 10, (4096,)))
        vs = TensorDataset(torch.randn(128, 3, 32, 32), torch.randint(0, 10, (128,)))
        return (BatchedTensorLoader(ds, 32, shuffle=True, drop_last=True, placement=PLACEMENT, device=self.device),
                BatchedTensorLoader(vs, 32, placement=PLACEMENT, device=self.device))
    def get_criterion(self, config): return nn.CrossEntropyLoss()
    def get_optimizer(self, config): return torch.optim.SGD(self.model.parameters(), lr=0.01, momentum=0.9)
c = C(Path("."), [Accuracy()], dev, client_name="p", engine_options=EngineOptions(cuda_graphs=True, amp_dtype=torch.bfloat16, channels_last=True, master_weights=True))
def cfg(r): return {"current_server_round": r, "local_steps": 8, "batch_size": 32}
params = c.get_parameters(cfg(0))
params = NDArrays([p.clone() for p in params])
def T():
    torch.cuda.synchronize(); return time.perf_counter()
acc = {}
def rec(k, t0):
    acc.setdefault(k, []).append((T() - t0) * 1e3)
for r in range(1, 9):
    t = T(); c.set_parameters(params, cfg(r), True); rec("set_parameters", t)
    t = T(); c.update_before_train(r); rec("update_before_train", t)
    t = T(); loss, met = c.train_by_steps(8, r); rec("train_by_steps(8)", t)
    t = T(); out = c.get_parameters(cfg(r)); rec("get_parameters", t)
    t = T(); agg = aggregate_results([(out, 100)], True); rec("aggregate(1 client)", t)
    params = agg
    t = T(); c.set_parameters(params, cfg(r), False); rec("set_parameters(eval)", t)
    t = T(); c.validate(); rec("validate(4)", t)
for k, v in acc.items():
    print(f"{k:24s} median {sorted(v[3:])[len(v[3:])//2]:7.3f} ms   (all: {' '.join(f'{x:.2f}' for x in v)})")

if os.environ.get("HOST_PROFILE", "0") == "1":
    import cProfile, pstats
    pr = cProfile.Profile()
    torch.cuda.synchronize()
    pr.enable()
    for r in range(9, 14):
        c.set_parameters(params, cfg(r), True)
        c.train_by_steps(8, r)
        c.validate()
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
