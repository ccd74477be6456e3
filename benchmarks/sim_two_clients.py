"""Single-GPU simulation of K ResNet-18 clients (sanity: FedAvg loss trajectory with the bench configuration)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ.setdefault("FL4H_LOG_LEVEL", "WARNING")
import torch
from torch import nn
from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.engine.data import BatchedTensorLoader
from fl4health_b200.engine.options import EngineOptions
from fl4health_b200.metrics import Accuracy
from fl4health_b200.metrics.metric_aggregation import evaluate_metrics_aggregation_fn, fit_metrics_aggregation_fn
from fl4health_b200.models import resnet18_cifar
from fl4health_b200.servers.base_server import FlServer
from fl4health_b200.servers.client_manager import SimpleClientManager
from fl4health_b200.simulation import run_simulation
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg
from fl4health_b200.utils.dataset import TensorDataset

K = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ROUNDS = int(sys.argv[2]) if len(sys.argv) > 2 else 13
def synthetic(n, seed):
    gen = torch.Generator().manual_seed(seed)
    t = torch.randint(0, 10, (n,), generator=gen)
    d = torch.randn(n, 3, 32, 32, generator=gen) * 0.5 + (t.float().view(-1, 1, 1, 1) - 4.5) * 0.1
    return TensorDataset(d, t)
class C(BasicClient):
    def __init__(self, *a, idx=0, **k):
        super().__init__(*a, **k); self.idx = idx
    def get_model(self, config):
        torch.manual_seed(1234); return resnet18_cifar()
    def get_data_loaders(self, config):
        return (BatchedTensorLoader(synthetic(4096, 100 + self.idx), 32, shuffle=True, drop_last=True, placement="device", device=self.device),
                BatchedTensorLoader(synthetic(128, 900 + self.idx), 32, placement="device", device=self.device))
    def get_criterion(self, config): return nn.CrossEntropyLoss()
    def get_optimizer(self, config): return torch.optim.SGD(self.model.parameters(), lr=0.01, momentum=0.9)
cfg = lambda r: {"current_server_round": r, "local_steps": 8, "batch_size": 32}
eng = EngineOptions(cuda_graphs=True, amp_dtype=torch.bfloat16, channels_last=True)
clients = [C(Path("."), [Accuracy()], torch.device("cuda:0"), client_name=f"c{i}", engine_options=eng, idx=i) for i in range(K)]
st = BasicFedAvg(min_fit_clients=K, min_evaluate_clients=K, min_available_clients=K, on_fit_config_fn=cfg, on_evaluate_config_fn=cfg,
                 fit_metrics_aggregation_fn=fit_metrics_aggregation_fn, evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn)
srv = FlServer(SimpleClientManager(), {"n_server_rounds": ROUNDS}, st, on_init_parameters_config_fn=cfg)
h = run_simulation(srv, clients, ROUNDS)
print("K", K, "val losses", [round(l, 3) for _, l in h.losses_distributed])
print("val acc", [round(a, 3) for _, a in h.metrics_distributed["val - prediction - accuracy"]])
