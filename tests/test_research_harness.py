"""The research harness (``research/``): every method builds and trains on a tiny task, artefacts have the layout the
selection tools read, hyper-parameter selection and held-out evaluation work, and real partition files are picked up."""

from __future__ import annotations

import json
from dataclasses import replace

import pytest
import torch

from fl4health_b200.utils.dataset import TensorDataset
from research.cifar10.model import conv_net
from research.cifar10.preprocess import partition, save_partitions
from research.harness import METHODS, TASKS, ExperimentSpec, evaluate_on_test, find_best_hp, run_experiment, sweep
from research.harness.selection import get_weighted_loss_from_server_log
from research.harness.servers import FullExchangeServer, PersonalServer, make_personal

CPU = torch.device("cpu")


def tiny(tmp_path, **overrides) -> ExperimentSpec:  # type: ignore[no-untyped-def]
    base = dict(task="synthetic", rounds=2, local_steps=2, batch_size=16, samples_per_client=120, beta_update_interval=2,
                artifact_dir=str(tmp_path / "out"), data_dir=str(tmp_path / "data"), task_kwargs={"n_clients": 2})
    base.update(overrides)
    return ExperimentSpec(**base)


def test_convnet_matches_the_benchmark_parameter_count() -> None:
    assert sum(p.numel() for p in conv_net().parameters()) == 8_465_034  # "8.47 M" in the reference's README
    assert conv_net()(torch.randn(2, 3, 32, 32)).shape == (2, 10)


@pytest.mark.parametrize("method", sorted(METHODS))
def test_every_method_runs_and_writes_the_run_layout(method: str, tmp_path) -> None:  # type: ignore[no-untyped-def]
    spec = tiny(tmp_path, method=method, checkpoint=method in ("fedavg", "ditto", "fenda", "apfl", "central"))
    (result,) = run_experiment(spec, CPU)
    run_dir = spec.hp_dir() / "Run1"
    assert len(result["losses_distributed"]) == 2 and all(torch.isfinite(torch.tensor(v)) for _, v in result["losses_distributed"])
    assert get_weighted_loss_from_server_log(run_dir) == pytest.approx(result["best_aggregated_loss"])
    assert json.loads((run_dir / "results.json").read_text())["spec"]["method"] == method
    if spec.checkpoint:
        n = 1 if method == "central" else 2
        assert all((run_dir / f"client_{i}_{w}_model.pkl").exists() for i in range(n) for w in ("best", "last"))
        assert (run_dir / "server_best_model.pkl").exists() == (method in ("fedavg", "central"))
        report = evaluate_on_test(spec.hp_dir(), CPU)
        assert 0.0 <= report["client_models_avg_accuracy"] <= 1.0
        assert ("server_model_avg_accuracy" in report) == (method in ("fedavg", "central"))


def test_sweep_and_best_hp_selection(tmp_path) -> None:  # type: ignore[no-untyped-def]
    spec = tiny(tmp_path, method="fedprox", runs=2, checkpoint=False)
    results = sweep(spec, {"lr": [0.0001, 0.05], "lam": [0.1]}, CPU)
    assert sorted(results) == ["lr_0.0001_lam_0.1", "lr_0.05_lam_0.1"] and all(len(runs) == 2 for runs in results.values())
    best_dir, best_loss = find_best_hp(spec.hp_dir().parent)
    means = {key: sum(r["best_aggregated_loss"] for r in runs) / 2 for key, runs in results.items()}
    assert best_dir.name == min(means, key=means.get) and best_loss == pytest.approx(min(means.values()))
    # the two repetitions of one setting use different training seeds
    a, b = results["lr_0.05_lam_0.1"]
    assert a["spec"]["seed"] + 1 == b["spec"]["seed"] and a["best_aggregated_loss"] != b["best_aggregated_loss"]
    with pytest.raises(FileNotFoundError):
        find_best_hp(tmp_path / "data" if (tmp_path / "data").exists() else tmp_path.parent / "nowhere-near")


def test_local_baseline_ignores_the_aggregate_after_round_one(tmp_path) -> None:  # type: ignore[no-untyped-def]
    from research.harness.experiment import build

    spec = tiny(tmp_path, method="local", checkpoint=False)
    _, clients, _ = build(spec, tmp_path, CPU)
    client = clients[0]
    client.setup_client({"current_server_round": 1, "batch_size": 16, "local_steps": 1})
    first = [torch.full_like(v, 0.25, dtype=torch.float32) for v in client.model.state_dict().values()]
    client.set_parameters(first, {"current_server_round": 1}, fitting_round=True)
    kept = [v.clone() for v in client.model.state_dict().values()]
    client.set_parameters([v * 0 for v in first], {"current_server_round": 2}, fitting_round=True)
    assert all(torch.equal(a, b) for a, b in zip(kept, client.model.state_dict().values()))


def test_central_baseline_pools_every_clients_data(tmp_path) -> None:  # type: ignore[no-untyped-def]
    from research.harness.experiment import build

    spec = tiny(tmp_path, method="central", checkpoint=False, task_kwargs={"n_clients": 3})
    _, clients, ctx = build(spec, tmp_path, CPU)
    assert len(clients) == 1
    pooled = clients[0].client_triple()
    singles = [ctx.task.client_data(i, spec) for i in range(3)]
    assert [len(p.data) for p in pooled] == [sum(len(s[k].data) for s in singles) for k in range(3)]


def test_personal_and_full_exchange_servers_track_the_best_loss() -> None:
    from fl4health_b200.servers.adaptive_constraint_servers.ditto_server import DittoServer

    cls = make_personal(DittoServer)
    assert issubclass(cls, DittoServer) and make_personal(cls) is cls and make_personal(PersonalServer) is PersonalServer
    tracker = PersonalServer.__new__(PersonalServer)
    for loss in (0.9, 1.2, 0.4, None, 0.6):
        tracker._track(loss)
    assert tracker.best_aggregated_loss == pytest.approx(0.4)
    assert FullExchangeServer.best_aggregated_loss is None


def test_dirichlet_partitions_share_label_marginals_and_are_used_by_the_task(tmp_path) -> None:  # type: ignore[no-untyped-def]
    gen = torch.Generator().manual_seed(0)
    make = lambda n: TensorDataset(torch.randn(n, 3, 32, 32, generator=gen), torch.randint(0, 10, (n,), generator=gen))  # noqa: E731
    parts = partition(make(3000), make(1000), make(1000), n_clients=3, beta=0.5, seed=11)
    assert 2950 <= sum(len(t.data) for t, _, _ in parts) <= 3000  # per-label allocations are floored, as in the reference
    for train, val, test in parts:
        p_train = torch.bincount(train.targets, minlength=10).float() / len(train.data)
        p_test = torch.bincount(test.targets, minlength=10).float() / len(test.data)
        assert (p_train - p_test).abs().max() < 0.1  # val / test follow the training allocation (fixed prior)
    spec = tiny(tmp_path, task="cifar10", method="fedavg", heterogeneity=0.5, data_seed=11, task_kwargs={"n_clients": 3, "hidden": 16})
    save_partitions(parts, tmp_path / "data" / "beta_0.5" / "seed_11")
    task = TASKS["cifar10"](n_clients=3, hidden=16)
    train, val, test = task.client_data(1, spec)
    assert torch.equal(train.data, parts[1][0].data) and torch.equal(test.targets, parts[1][2].targets)
    # without files the task falls back to seeded synthetic shards (deterministic per client and seed)
    fallback = task.client_data(1, replace(spec, data_seed=12))
    again = task.client_data(1, replace(spec, data_seed=12))
    assert fallback[0].data.shape[1:] == (3, 32, 32) and torch.equal(fallback[0].data, again[0].data)


@pytest.mark.parametrize("name, kwargs, shape", [("ag_news", {"n_clients": 2}, (32,)), ("rxrx1", {"n_clients": 2, "class_num": 5, "image_size": 32}, (3, 32, 32)),
                                                 ("fed_heart_disease", {}, (13,))])
def test_other_tasks_build_models_and_data(name: str, kwargs: dict, shape: tuple, tmp_path) -> None:  # type: ignore[no-untyped-def]
    task = TASKS[name](**kwargs)
    spec = tiny(tmp_path, task=name, samples_per_client=40)
    train, val, test = task.client_data(0, spec)
    assert train.data.shape[1:] == shape and len(val.data) > 0 and len(test.data) > 0
    logits = task.plain()(train.data[:4])
    assert logits.shape == (4, task.class_num)
    features = task.features()(train.data[:4])
    assert task.parallel_head()(features, features).shape == (4, task.class_num)


def test_command_line_round_trip(tmp_path, capsys) -> None:  # type: ignore[no-untyped-def]
    from research.run import main

    common = ["--task", "synthetic", "--method", "mr_mtl", "--rounds", "1", "--local-steps", "1", "--samples-per-client", "80", "--no-checkpoint",
              "--artifact-dir", str(tmp_path / "o"), "--device", "cpu", "--task-kwargs", "n_clients=2"]
    main(["sweep", *common, "--grid", "lam=0.5,2.0"])
    printed = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert sorted(printed) == ["lr_0.01_lam_0.5", "lr_0.01_lam_2.0"]
    folder, _ = main(["best", "--dir", str(tmp_path / "o" / "synthetic" / "mr_mtl")])
    assert folder.name in printed


@pytest.mark.parametrize("method, n_clients", [("fedavg", 2), ("ditto", 4)])
def test_spmd_run_matches_the_in_process_simulation(method: str, n_clients: int, tmp_path) -> None:  # type: ignore[no-untyped-def]
    """``torchrun … -m research.run train --spmd`` (gloo, 2 ranks; 1 or 2 clients per rank) reproduces the simulation."""
    import os
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parents[1]
    common = ["--task", "synthetic", "--method", method, "--rounds", "2", "--local-steps", "2", "--samples-per-client", "100", "--no-checkpoint",
              "--device", "cpu", "--task-kwargs", f"n_clients={n_clients}"]
    from research.run import main

    (reference,) = main(["train", *common, "--artifact-dir", str(tmp_path / "sim")])
    port = 29700 + (os.getpid() + n_clients) % 200
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", str(port),
           "-m", "research.run", "train", *common, "--artifact-dir", str(tmp_path / "spmd"), "--spmd"]
    env = {**os.environ, "FL4H_LOG_LEVEL": "ERROR", "PYTHONPATH": str(root)}
    proc = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600, cwd=root)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [line for line in proc.stdout.splitlines() if line.startswith("[{")]
    assert len(lines) == 1  # rank 0 reports
    assert json.loads(lines[0])[0]["best_aggregated_loss"] == pytest.approx(reference["best_aggregated_loss"], rel=1e-5)
