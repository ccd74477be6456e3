"""More clients than ranks (parallel/spmd_multi.py): 4 clients over 2 gloo ranks — evenly and unevenly split — must
reproduce the single-process federation of the same 4 clients."""

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


def _launch(tmp_path: Path, strategy: str, split: str, port: int, **extra_env: str) -> dict:
    out = tmp_path / f"{strategy}.json"
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", FL4H_LOG_LEVEL="ERROR", **extra_env)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "tests" / "spmd_multi_worker.py"), str(out), strategy, split]
    proc = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    return json.loads(out.read_text())


def _local(strategy_name: str) -> dict:
    set_all_random_seeds(42)
    common = dict(min_fit_clients=4, min_evaluate_clients=4, min_available_clients=4, on_fit_config_fn=fit_config_fn(),
                  on_evaluate_config_fn=fit_config_fn(), fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
                  evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn)
    if strategy_name == "fedadam":
        from fl4health_b200.common.typing import ndarrays_to_parameters
        from fl4health_b200.models import Net
        from fl4health_b200.parallel.arena import attach_arena

        torch.manual_seed(1234)
        arena = attach_arena(Net(), "cpu", with_grad=False)
        strategy = FedAdam(initial_parameters=ndarrays_to_parameters(arena.ndarrays()), eta=0.05, **common)
    else:
        strategy = BasicFedAvg(**common)
    server = FlServer(SimpleClientManager(), {"n_server_rounds": 2}, strategy, on_init_parameters_config_fn=fit_config_fn())
    clients = make_clients(4)
    history = run_simulation(server, clients, num_rounds=2)
    state = {k: v.detach().cpu().double().sum().item() for k, v in clients[0].model.state_dict().items()}
    return {"losses": history.losses_distributed, "state": state}


@pytest.mark.parametrize("strategy,split,port", [("fedavg", "2,2", 29631), ("fedavg", "3,1", 29632), ("fedadam", "2,2", 29633)])
def test_multi_client_per_rank_matches_single_process(tmp_path: Path, strategy: str, split: str, port: int) -> None:
    spmd = _launch(tmp_path, strategy, split, port)
    assert spmd["clients"] == 4
    local = _local(strategy)
    # FedAdam divides by sqrt(v) of a first-round second moment: the different (but fixed) summation order of the
    # two-level reduce — per-rank partial sums, then across ranks — is amplified there; FedAvg matches to rounding
    rel = 2e-3 if strategy == "fedadam" else 0.0
    for (r1, l1), (r2, l2) in zip(spmd["losses"], local["losses"]):
        assert r1 == r2 and abs(l1 - l2) < 1e-5 + rel * abs(l2), (spmd["losses"], local["losses"])
    for key, value in local["state"].items():
        assert abs(spmd["state"][key] - value) < (1e-4 + 5 * rel) * max(1.0, abs(value)), key


@pytest.mark.parametrize("scenario,port", [("fedprox_example", 29641), ("scaffold_example", 29642)])
def test_packed_payload_scenarios_with_two_clients_per_rank(tmp_path: Path, scenario: str, port: int) -> None:
    """Payloads with packed side information (adaptive loss weight, control variates) through the multi-client path."""
    summaries = _run_scenario_spmd(scenario, port, "--clients-per-rank", "2")
    assert summaries and all(len(s["losses"]) == 2 and all(loss == loss for _, loss in s["losses"]) for s in summaries)


def _run_scenario_spmd(scenario: str, port: int, *extra: str) -> list[dict]:
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", FL4H_LOG_LEVEL="ERROR")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), "-m", "examples.run", scenario, "--spmd", *extra, "--rounds", "2", "--device", "cpu"]
    proc = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    return [json.loads(line) for line in proc.stdout.splitlines() if line.startswith(chr(123) + chr(34) + "scenario")]


@pytest.mark.parametrize("scenario,port", [("scaffold_example", 29661), ("dynamic_layer_exchange_example", 29662),
                                           ("client_level_dp_example", 29663), ("fedper_example", 29664), ("fedbn_example", 29665)])
def test_strategies_that_need_whole_payloads_run_in_spmd_and_match_simulation(scenario: str, port: int) -> None:
    """SCAFFOLD (packed variates), dynamic layer exchange and client-level DP either reduce packed side payloads or
    materialise every client's payload on every rank (the owner of a payload used to skip that broadcast: deadlock);
    FedPer / FedBN exchange a named SUBSET of the arena, which rides the whole-arena reduction and is pulled as a few
    range copies."""
    from examples.run import main

    spmd = _run_scenario_spmd(scenario, port)
    local = main([scenario, "--rounds", "2", "--clients", "2", "--device", "cpu"])
    assert len(spmd) == 1  # rank 0 reports (every rank holds the same history)
    for summary in spmd:
        for (r1, l1), (r2, l2) in zip(summary["losses"], local["losses"]):
            assert r1 == r2 and abs(l1 - l2) < 1e-5, (summary["losses"], local["losses"])


def test_fedpm_votes_on_bit_packed_masks_across_ranks() -> None:
    """FedPM with one client per rank: the masks cross ranks as packed words in one all-gather (``strategies/fedpm.py``);
    two real processes over gloo, and the posterior keeps producing finite, changing losses."""
    (summary,) = _run_scenario_spmd("fedpm_example", 29681)
    losses = [loss for _, loss in summary["losses"]]
    assert len(losses) == 2 and all(loss == loss for loss in losses) and losses[0] != losses[1]


def test_partial_participation_with_idle_ranks(tmp_path: Path) -> None:
    """Half of six clients (hosted 1 + 5) are sampled each round for five rounds: rounds in which a rank has no selected
    client — including before it has ever seen a payload — must neither deadlock nor change the collective sequence."""
    import math

    result = _launch(tmp_path, "fedavg", "1,5", 29671, FL4H_TEST_FRACTION="0.5", FL4H_TEST_ROUNDS="5")
    assert result["clients"] == 6 and len(result["losses"]) == 5
    assert all(math.isfinite(loss) for _, loss in result["losses"])


def test_one_client_per_rank_with_partial_participation(tmp_path: Path) -> None:
    """One client per rank, half of them sampled per round: the rank that sits a round out contributes weight 0 from its
    own arena (or, before it has ever produced a payload, everybody agrees on the packed route) instead of raising."""
    import math

    result = _launch(tmp_path, "fedavg", "1,1", 29672, FL4H_TEST_FRACTION="0.5", FL4H_TEST_ROUNDS="6")
    assert result["clients"] == 2 and len(result["losses"]) == 6
    assert all(math.isfinite(loss) for _, loss in result["losses"])


def test_clients_without_metrics_survive_the_cached_metadata_schema() -> None:
    """A client with no metrics sends an EMPTY metrics dict: from round 2 on the numbers-only metadata exchange has to
    rebuild that (empty) dict instead of dropping the key."""
    (summary,) = _run_scenario_spmd("ae_example", 29673)
    assert len(summary["losses"]) == 2 and summary["metrics"] == {}
