"""Clients spanning several ranks (``fl4health_b200.parallel.client_group``): gradient averaging equals full-batch
training, shards are disjoint, and a 2-rank client behaves like the single-process client on the union of its data."""

from __future__ import annotations

import json
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

from fl4health_b200.parallel.client_group import ClientGroup, ReplicatedClientMixin, average_gradients, shard_dataset
from fl4health_b200.utils.dataset import TensorDataset

ROOT = Path(__file__).resolve().parents[1]


def test_shards_are_disjoint_equal_and_cover_the_data() -> None:
    data = TensorDataset(torch.arange(23).float().view(-1, 1), torch.arange(23))
    groups = [ClientGroup(0, r, 3, None) for r in range(3)]
    shards = [shard_dataset(data, g, seed=4) for g in groups]
    assert [len(s.data) for s in shards] == [7, 7, 7]
    merged = torch.cat([s.targets for s in shards])
    assert len(set(merged.tolist())) == 21  # disjoint; 23 % 3 samples dropped so that replicas stay in lock-step
    assert shard_dataset(data, ClientGroup(0, 0, 1, None)) is data
    assert all(torch.equal(a.targets, b.targets) for a, b in zip(shards, [shard_dataset(data, g, seed=4) for g in groups]))
    with pytest.raises(ValueError):
        ClientGroup.from_world(0, 6, 4)
    solo = ClientGroup.from_world(3, 8, 1)
    assert (solo.client_index, solo.group_rank, solo.process_group) == (3, 0, None)


def test_single_rank_group_leaves_gradients_alone() -> None:
    model = torch.nn.Linear(3, 2)
    model(torch.randn(4, 3)).sum().backward()
    before = [p.grad.clone() for p in model.parameters()]
    average_gradients(model, ClientGroup(0, 0, 1, None))
    assert all(torch.equal(a, p.grad) for a, p in zip(before, model.parameters()))

    class Base:
        calls = 0

        def transform_gradients(self, losses) -> None:  # type: ignore[no-untyped-def]
            Base.calls += 1

    class Client(ReplicatedClientMixin, Base):
        pass

    client = Client()
    client.model = model
    client.transform_gradients(None)  # composes through super() even without a group
    assert Base.calls == 1


def test_strategies_that_count_participants_refuse_replicated_clients() -> None:
    from fl4health_b200.parallel.client_group import require_replica_safe_strategy
    from fl4health_b200.strategies.basic_fedavg import BasicFedAvg
    from fl4health_b200.strategies.client_dp_fedavgm import ClientLevelDPFedAvgM
    from fl4health_b200.strategies.scaffold import Scaffold
    from fl4health_b200.common.typing import ndarrays_to_parameters

    require_replica_safe_strategy(BasicFedAvg(), 2)
    dp = ClientLevelDPFedAvgM(initial_parameters=ndarrays_to_parameters([torch.zeros(2)]))
    require_replica_safe_strategy(dp, 1)  # one rank per client: anything goes
    scaffold = Scaffold(initial_parameters=ndarrays_to_parameters([torch.zeros(2)]), model=torch.nn.Linear(1, 1, bias=False))
    for strategy in (dp, scaffold):
        with pytest.raises(ValueError, match="one rank per client"):
            require_replica_safe_strategy(strategy, 2)


def _launch(tmp_path: Path, world: int, group_size: int, port: int, **extra_env: str) -> list[dict]:
    out = tmp_path / f"w{world}g{group_size}{'z' if extra_env else ''}"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "tests" / "client_group_worker.py"), str(out), str(group_size)]
    proc = subprocess.run(cmd, env={**os.environ, "FL4H_LOG_LEVEL": "ERROR", **extra_env}, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-3000:]
    return [json.loads(Path(f"{out}.rank{r}").read_text()) for r in range(world)]


def test_two_rank_client_matches_the_single_rank_client_on_the_union_batch(tmp_path: Path) -> None:
    base = 29900 + os.getpid() % 50
    (solo,) = _launch(tmp_path, world=1, group_size=1, port=base)
    replicas = _launch(tmp_path, world=2, group_size=2, port=base + 1)
    # both replicas hold the same parameters before every aggregate (they took identical, averaged steps) ...
    assert all(r["grad_error"] < 1e-6 for r in replicas)  # averaged shard gradients == union-batch gradient
    for key in ("1", "2"):
        assert replicas[0]["pre_aggregate"][key] == pytest.approx(replicas[1]["pre_aggregate"][key], rel=1e-9)
    # ... each trains on half of the client's data, and the federated losses are finite and improving
    assert [r["train_samples"] for r in replicas] == [64, 64] and solo["train_samples"] == 128
    losses = [v for _, v in replicas[0]["losses"]]
    assert losses[1] < losses[0] and all(torch.isfinite(torch.tensor(v)) for v in losses)
    assert replicas[0]["losses"] == replicas[1]["losses"]


def test_four_ranks_two_clients_of_two_replicas(tmp_path: Path) -> None:
    results = _launch(tmp_path, world=4, group_size=2, port=29960 + os.getpid() % 30)
    assert [r["client"] for r in results] == [0, 0, 1, 1]
    for a, b in ((0, 1), (2, 3)):  # replicas of one client agree; the two clients (different data) do not
        assert results[a]["pre_aggregate"]["1"] == pytest.approx(results[b]["pre_aggregate"]["1"], rel=1e-9)
    assert results[0]["pre_aggregate"]["1"] != pytest.approx(results[2]["pre_aggregate"]["1"], rel=1e-6)
    assert all(r["losses"] == results[0]["losses"] for r in results)


@pytest.mark.parametrize("variant", ["ditto", "apfl"])
def test_replicas_of_multi_optimizer_clients_stay_identical(tmp_path: Path, variant: str) -> None:
    """Clients that write their own ``train_step`` around several optimizers (Ditto: exchanged twin + personal model;
    APFL: global + local sub-models) never go through ``transform_gradients``: the group average hangs on
    ``optimizer.step()`` instead, and every model of the client -- including the ones that are never exchanged, which
    only the averaging keeps in sync -- is the same on both replicas before each aggregate."""
    replicas = _launch(tmp_path, world=2, group_size=2, port=29750 + (os.getpid() + len(variant)) % 40, FL4H_TEST_VARIANT=variant)
    for key in ("1", "2"):
        assert replicas[0]["pre_aggregate"][key] == pytest.approx(replicas[1]["pre_aggregate"][key], rel=1e-9)
    assert replicas[0]["pre_aggregate"]["1"] != pytest.approx(replicas[0]["pre_aggregate"]["2"], rel=1e-6)  # and they do train
    assert replicas[0]["losses"] == replicas[1]["losses"]


def test_partition_is_balanced_and_deterministic() -> None:
    from fl4health_b200.parallel.client_group import partition_parameters

    params = [torch.nn.Parameter(torch.zeros(n)) for n in (100, 10, 90, 50, 45, 5)]
    table = partition_parameters(params, 2)
    assert sorted(i for owned in table for i in owned) == list(range(6))
    loads = [sum(params[i].numel() for i in owned) for owned in table]
    assert abs(loads[0] - loads[1]) <= 10 and table == partition_parameters(params, 2)
    assert partition_parameters(params[:1], 3) == [[0], [], []]


def test_zero1_matches_replicated_training_with_half_the_optimizer_state(tmp_path: Path) -> None:
    base = 29850 + os.getpid() % 40
    plain = _launch(tmp_path, world=2, group_size=2, port=base)
    sharded = _launch(tmp_path, world=2, group_size=2, port=base + 1, FL4H_TEST_ZERO1="1")
    # same trajectory: ownership only decides WHO computes an update, not its value
    for key in ("1", "2"):
        assert sharded[0]["pre_aggregate"][key] == pytest.approx(plain[0]["pre_aggregate"][key], rel=1e-6)
        assert sharded[0]["pre_aggregate"][key] == pytest.approx(sharded[1]["pre_aggregate"][key], rel=1e-9)
    assert sharded[0]["final"] == pytest.approx(plain[0]["final"], rel=1e-5, abs=1e-7)
    assert [v for _, v in sharded[0]["losses"]] == pytest.approx([v for _, v in plain[0]["losses"]], rel=1e-6)
    # momentum buffers: together the shards hold exactly one copy of the state, tensor-granular, so this tiny model (one 2560-element matrix) splits unevenly
    shards, total = [r["optimizer_state_elements"] for r in sharded], sharded[0]["trainable_elements"]
    assert sum(shards) == total and 0 < min(shards) and max(shards) < total and sharded[0]["optimizer_class"] == "SGD"


@pytest.mark.parametrize("scenario", ["fedllm_example", "ditto_example"])
def test_examples_run_with_two_ranks_per_client(scenario: str) -> None:
    """``examples.run --spmd --ranks-per-client 2``: the LLM example the reference shards with DeepSpeed, and a client
    that trains two models per step (both are averaged over the group)."""
    port = 29800 + (os.getpid() + len(scenario)) % 40
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", str(port),
           "-m", "examples.run", scenario, "--spmd", "--ranks-per-client", "2", "--device", "cpu", "--rounds", "2"]
    proc = subprocess.run(cmd, env={**os.environ, "FL4H_LOG_LEVEL": "ERROR"}, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-3000:]
    (line,) = [ln for ln in proc.stdout.splitlines() if ln.startswith("{")]
    losses = [v for _, v in json.loads(line)["losses"]]
    assert len(losses) == 2 and all(torch.isfinite(torch.tensor(v)) for v in losses)
