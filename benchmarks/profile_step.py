"""Steady-state profile harness for the flagship local step (ResNet-18, batch 32, bf16, CUDA graph).

Runs warm-up (eager + capture) and then brackets a few graph replays with cudaProfilerStart/Stop so that
``ncu --profile-from-start off`` sees only steady-state kernels.  Also prints device-timed phase costs (CUDA events).

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/launches.csv python benchmarks/profile_step.py
"""

from __future__ import annotations

import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ.setdefault("FL4H_LOG_LEVEL", "WARNING")

import torch  # noqa: E402
from torch import nn  # noqa: E402

from fl4health_b200.clients.basic_client import BasicClient  # noqa: E402
from fl4health_b200.engine.data import BatchedTensorLoader  # noqa: E402
from fl4health_b200.engine.options import EngineOptions  # noqa: E402
from fl4health_b200.metrics import Accuracy  # noqa: E402
from fl4health_b200.models import resnet18_cifar  # noqa: E402
from fl4health_b200.utils.dataset import TensorDataset  # noqa: E402

BS = int(os.environ.get("BS", "32"))


class Client(BasicClient):
    def get_model(self, config):  # noqa: ANN001, ANN201
        return resnet18_cifar()

    def get_data_loaders(self, config):  # noqa: ANN001, ANN201
        ds = TensorDataset(torch.randn(2048, 3, 32, 32), torch.randint(0, 10, (2048,)))
        return (BatchedTensorLoader(ds, BS, shuffle=True, drop_last=True, placement="device", device=self.device),
                BatchedTensorLoader(ds, BS, placement="device", device=self.device))

    def get_criterion(self, config):  # noqa: ANN001, ANN201
        return nn.CrossEntropyLoss()

    def get_optimizer(self, config):  # noqa: ANN001, ANN201
        return torch.optim.SGD(self.model.parameters(), lr=0.01, momentum=0.9)


def timed(fn, iters: int) -> float:  # noqa: ANN001
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(iters):
        fn()
    end.record()
    torch.cuda.synchronize()
    return start.elapsed_time(end) / iters


def main() -> None:
    device = torch.device("cuda:0")
    torch.backends.cudnn.benchmark = True
    graphs = os.environ.get("GRAPHS", "1") == "1"
    engine = EngineOptions(cuda_graphs=graphs, amp_dtype=torch.bfloat16, channels_last=True,
                           master_weights=os.environ.get("MASTER", "1") == "1")
    client = Client(Path("."), [Accuracy()], device, client_name="prof", engine_options=engine)
    cfg = {"current_server_round": 1, "local_steps": 8, "batch_size": BS}
    client.setup_client(cfg)
    client.model.train()
    it = iter(client.train_loader)
    x, y = next(it)
    x, y = client._prepare_batch(x, y)

    def train_step() -> None:
        client._run_train_unit(x, y)

    for _ in range(6):
        train_step()
    torch.cuda.synchronize()
    ms_train = timed(train_step, 50)
    t0 = time.perf_counter()
    for _ in range(50):
        train_step()
    host_issue_ms = (time.perf_counter() - t0) * 1000 / 50
    torch.cuda.synchronize()

    client.model.eval()

    def val_step() -> None:
        client._run_val_unit(x, y, client.val_loss_meter, client.val_metric_manager)

    for _ in range(6):
        val_step()
    ms_val = timed(val_step, 50)

    def next_batch() -> None:
        client._prepare_batch(*client._next_train_batch())

    ms_batch = timed(next_batch, 50)
    arena_flat = client.model.parameters().__next__()
    print(f"batch={BS} graphs={graphs} train_step_ms={ms_train:.3f} (host issue {host_issue_ms:.3f}) val_step_ms={ms_val:.3f} "
          f"next_batch_ms={ms_batch:.3f} replays={getattr(client._train_runner, 'replays', 0)}")

    client.model.train()
    torch.cuda.cudart().cudaProfilerStart()
    train_step()
    train_step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
