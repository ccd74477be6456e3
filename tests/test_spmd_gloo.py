"""SPMD runtime on CPU (gloo, 2 processes) must reproduce the single-process simulation bit-for-bit-ish."""

import json
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

from fl4health_b200.metrics.metric_aggregation import evaluate_metrics_aggregation_fn, fit_metrics_aggregation_fn
from fl4health_b200.servers.base_server import FlServer
from fl4health_b200.servers.client_manager import SimpleClientManager
from fl4health_b200.simulation import run_simulation
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg
from fl4health_b200.strategies.fedopt import FedAdam
from fl4health_b200.utils.random import set_all_random_seeds
from tests.helpers import fit_config_fn, make_clients

ROOT = Path(__file__).resolve().parent.parent


def _launch(tmp_path: Path, strategy: str, port: int, **extra_env: str) -> dict:
    out = tmp_path / f"{strategy}.json"
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", FL4H_LOG_LEVEL="ERROR", **extra_env)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(ROOT / "tests" / "spmd_worker.py"), str(out), strategy]
    proc = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    return json.loads(out.read_text())


def _local(strategy_name: str) -> dict:
    set_all_random_seeds(42)
    common = dict(min_fit_clients=2, min_evaluate_clients=2, min_available_clients=2,
                  on_fit_config_fn=fit_config_fn(), on_evaluate_config_fn=fit_config_fn(),
                  fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
                  evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn)
    if strategy_name == "fedadam":
        from fl4health_b200.common.typing import ndarrays_to_parameters
        from fl4health_b200.models import Net
        from fl4health_b200.parallel.arena import attach_arena

        torch.manual_seed(1234)
        template = Net()
        arena = attach_arena(template, "cpu", with_grad=False)
        strategy = FedAdam(initial_parameters=ndarrays_to_parameters(arena.ndarrays()), eta=0.05, **common)
    else:
        strategy = BasicFedAvg(**common)
    server = FlServer(SimpleClientManager(), {"n_server_rounds": 2}, strategy,
                      on_init_parameters_config_fn=fit_config_fn())
    clients = make_clients(2)
    # SPMD client ids are rank-based; give the local run the same names so sampling/order match
    history = run_simulation(server, clients, num_rounds=2)
    state = {k: v.detach().cpu().double().sum().item() for k, v in clients[0].model.state_dict().items()}
    return {"losses": history.losses_distributed, "state": state}


@pytest.mark.parametrize("strategy,port,mailbox", [("fedavg", 29611, "1"), ("fedadam", 29612, "1"), ("fedavg", 29613, "0")])
def test_spmd_matches_single_process(tmp_path: Path, strategy: str, port: int, mailbox: str) -> None:
    """``mailbox``: per-round metadata through the shared-memory mailbox (default) or, with ``FL4H_SHM_MAILBOX=0``,
    through the fixed-size float64 all-gather — both must reproduce the single-process federation."""
    spmd = _launch(tmp_path, strategy, port, FL4H_SHM_MAILBOX=mailbox)
    from fl4health_b200.runtime.mailbox import load_runtime

    assert spmd["mailbox"] == (mailbox == "1" and load_runtime() is not None)
    local = _local(strategy)
    for (r1, l1), (r2, l2) in zip(spmd["losses"], local["losses"]):
        assert r1 == r2 and abs(l1 - l2) < 1e-5, (spmd["losses"], local["losses"])
    for key, value in local["state"].items():
        assert abs(spmd["state"][key] - value) < 1e-4 * max(1.0, abs(value)), key
