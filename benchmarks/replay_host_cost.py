"""Host duration of every cudaGraphLaunch (CUDAGraph.replay) inside real FL rounds, no profiler attached: is the
first replay of a round slower than the rest?"""
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
WARM, ROUNDS = 4, 10


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
import time
durations = []
_orig = torch.cuda.CUDAGraph.replay


def timed_replay(self):
    t0 = time.perf_counter()
    _orig(self)
    durations.append((time.perf_counter() - t0) * 1e6)


torch.cuda.CUDAGraph.replay = timed_replay
marks = []
server.round_end_hooks = [lambda r: marks.append(len(durations))]
t0 = time.perf_counter()
server.fit(num_rounds=WARM + ROUNDS)
torch.cuda.synchronize()
for r in range(WARM, WARM + ROUNDS):
    chunk = durations[marks[r - 1]:marks[r]]
    print(f"round {r+1}: {len(chunk)} replays, host us: " + " ".join(f"{d:.0f}" for d in chunk))
