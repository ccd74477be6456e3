"""The benchmark's reference arm: the flwr shim in ``baseline/stubs`` and the unmodified reference package driven
through it (CPU plumbing run; the GPU numbers come from ``bench.py --impl reference`` on the box)."""

from __future__ import annotations

import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
STUBS = ROOT / "baseline" / "stubs"


def _run(code: str) -> str:
    env = dict(os.environ, PYTHONPATH=f"{STUBS}:{ROOT / 'baseline' / '_ref'}")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300, cwd="/tmp")
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout


def test_shim_is_self_contained_and_round_trips_parameters() -> None:
    out = _run(
        "import sys, numpy as np, flwr\n"
        "from flwr.common import ndarrays_to_parameters, parameters_to_ndarrays\n"
        "from flwr.server.strategy.aggregate import aggregate, weighted_loss_avg\n"
        "a=[np.arange(6,dtype=np.float32).reshape(2,3), np.array(3,dtype=np.int64)]\n"
        "b=parameters_to_ndarrays(ndarrays_to_parameters(a))\n"
        "assert all((x==y).all() and x.dtype==y.dtype for x,y in zip(a,b))\n"
        "m=aggregate([([np.ones(3)],1),([np.zeros(3)],3)])\n"
        "assert np.allclose(m[0],0.25) and abs(weighted_loss_avg([(1,4.0),(3,0.0)])-1.0)<1e-12\n"
        "assert not any(k.startswith('fl4health_b200') for k in sys.modules)\n"
        "print('ok')\n"
    )
    assert out.strip() == "ok"


def test_shim_server_and_clients_over_tcp() -> None:
    """FedAvg of two NumPyClients through start_server/start_client (threads, localhost TCP)."""
    out = _run(
        "import threading, socket, numpy as np, flwr\n"
        "from flwr.client import NumPyClient, start_client\n"
        "from flwr.server import start_server, ServerConfig\n"
        "from flwr.server.strategy import FedAvg\n"
        "from flwr.common import ndarrays_to_parameters\n"
        "class C(NumPyClient):\n"
        "    def __init__(s,v,n): s.v=v; s.n=n\n"
        "    def fit(s,p,c): return [p[0]+s.v], s.n, {}\n"
        "    def evaluate(s,p,c): return float(p[0].sum()), s.n, {}\n"
        "sock=socket.socket(); sock.bind(('127.0.0.1',0)); port=sock.getsockname()[1]; sock.close()\n"
        "ts=[threading.Thread(target=start_client,kwargs=dict(server_address=f'127.0.0.1:{port}',client=C(v,n).to_client(),cid=str(v))) for v,n in ((1.0,1),(5.0,3))]\n"
        "[t.start() for t in ts]\n"
        "h=start_server(server_address=f'127.0.0.1:{port}',config=ServerConfig(num_rounds=2),strategy=FedAvg(initial_parameters=ndarrays_to_parameters([np.zeros(2)])))\n"
        "[t.join(30) for t in ts]\n"
        "print(h.losses_distributed)\n"
    )
    # each round adds the weighted mean (1*1+5*3)/4 = 4 to both entries -> loss = sum = 8, then 16
    assert out.strip() == "[(1, 8.0), (2, 16.0)]"


@pytest.mark.skipif(not (ROOT / "baseline" / "_ref" / "fl4health").exists() and not Path("/root/reference/fl4health").exists(),
                    reason="reference package not present")
def test_reference_arm_runs_the_unmodified_reference_on_cpu() -> None:
    cmd = [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--device", "cpu", "--steps", "1", "--warmup", "1",
           "--local-steps", "1", "--val-batches", "1", "--train-samples", "64", "--skip-e2e"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and "unavailable" not in line
    assert line["metric"] == "fl_rounds_per_sec_cifar10_resnet18_fedavg" and line["ms_per_step"] > 0
    assert line["config"]["exchange_payload_bytes"] == 44734408  # ResNet-18/CIFAR state_dict, fp32 + int64 counters
    assert np.isfinite(line["final_val_loss"])
    # the copy is unmodified
    if Path("/root/reference/fl4health").exists():
        diff = subprocess.run(["diff", "-r", "-q", "-x", "__pycache__", "/root/reference/fl4health",
                               str(ROOT / "baseline" / "_ref" / "fl4health")], capture_output=True, text=True)
        assert diff.returncode == 0, diff.stdout[:500]


def test_shim_accept_loop_survives_abandoned_connections_and_a_connect_storm() -> None:
    """Eight clients connect within the same instant (one rank per GPU does) and some connection attempts are dropped
    half-way through the authentication handshake: the accept loop has to keep accepting.  (It used to die with the
    first EOFError, after which the remaining clients were never registered and an 8-GPU reference run hung.)"""
    out = _run(
        "import threading, socket, time, numpy as np, flwr\n"
        "from flwr.client import NumPyClient, start_client\n"
        "from flwr.server import start_server, ServerConfig\n"
        "from flwr.server.strategy import FedAvg\n"
        "from flwr.common import ndarrays_to_parameters\n"
        "class C(NumPyClient):\n"
        "    def fit(s,p,c): return [p[0]+1.0], 1, {}\n"
        "    def evaluate(s,p,c): return float(p[0].sum()), 1, {}\n"
        "sock=socket.socket(); sock.bind(('127.0.0.1',0)); port=sock.getsockname()[1]; sock.close()\n"
        "def rude():\n"
        "    for _ in range(200):\n"
        "        try:\n"
        "            s=socket.create_connection(('127.0.0.1',port),timeout=1); s.close(); return\n"
        "        except OSError: time.sleep(0.05)\n"
        "K=8\n"
        "ts=[threading.Thread(target=rude) for _ in range(4)]\n"
        "ts+=[threading.Thread(target=start_client,kwargs=dict(server_address=f'127.0.0.1:{port}',client=C().to_client(),cid=str(i))) for i in range(K)]\n"
        "[t.start() for t in ts]\n"
        "h=start_server(server_address=f'127.0.0.1:{port}',config=ServerConfig(num_rounds=2),"
        "strategy=FedAvg(min_fit_clients=K,min_evaluate_clients=K,min_available_clients=K,initial_parameters=ndarrays_to_parameters([np.zeros(2)])))\n"
        "[t.join(30) for t in ts]\n"
        "print(h.losses_distributed)\n"
    )
    assert out.strip() == "[(1, 2.0), (2, 4.0)]"
