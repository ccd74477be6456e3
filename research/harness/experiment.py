"""Run one experiment (task × method × hyper-parameters × repeated runs) or a sweep over a grid.

On-disk layout (the one ``find_best_hp`` / ``evaluate_on_test`` read; same shape as the reference's sweep artefacts,
``research/cifar10/*/run_hp_sweep.sh`` + ``run_fold_experiment.slrm``)::

    <artifact_dir>/<task>/<method>/<hp key, e.g. lr_0.01_lam_1.0>/Run1/
        server.out                      # progress lines; LAST line = best weighted validation loss of the run
        results.json                    # spec, per-round losses / metrics, wall-clock
        client_<i>_best_model.pkl       # per-client checkpoints (post-aggregation, best validation loss / latest)
        client_<i>_last_model.pkl
        server_best_model.pkl           # only for full-exchange methods
        server_last_model.pkl

Repeated runs differ in the training seed (and, when ``vary_data`` is set, in the partition seed).  Execution is an
in-process simulation on one device by default; under ``torchrun`` (``--spmd``) every rank hosts a slice of the clients
and aggregation runs through the fused NVLink collectives (``fl4health_b200.parallel``).
"""

from __future__ import annotations

import itertools
import json
import time
from dataclasses import asdict, dataclass, replace
from logging import INFO
from pathlib import Path
from typing import Any

import torch

from fl4health_b200.common.logger import log


@dataclass
class ExperimentSpec:
    task: str = "cifar10"
    method: str = "fedavg"
    # federation
    rounds: int = 10
    local_steps: int = 50
    local_epochs: int | None = None
    batch_size: int = 32
    # optimisation
    optimizer: str = "sgd"
    lr: float = 0.01
    momentum: float = 0.9
    weight_decay: float = 0.001
    server_lr: float = 0.1          # FedAdam / FedYogi η
    server_lr_scaffold: float = 1.0
    alpha_lr: float = 0.01          # APFL mixing-parameter learning rate
    # method penalties
    lam: float = 1.0                # FedProx / Ditto / MR-MTL λ, MOON / PerFCL μ
    lam_delta: float = 0.1
    lam_patience: int = 5
    mmd_weight: float = 1.0
    beta_update_interval: int = 20
    exchange_fraction: float = 0.5
    # data
    data_dir: str = "research_data"
    heterogeneity: float = 0.5      # Dirichlet β of the partition
    samples_per_client: int = 2000
    data_seed: int = 2021
    vary_data: bool = False
    # bookkeeping
    seed: int = 2021
    runs: int = 1
    artifact_dir: str = "research_out"
    checkpoint: bool = True
    evaluate_test_each_round: bool = False
    task_kwargs: dict[str, Any] | None = None

    def hp_key(self) -> str:
        parts = [f"lr_{self.lr}"]
        if self.method in PENALISED:
            parts.append(f"lam_{self.lam}")
        if "mmd" in self.method:
            parts.append(f"mmd_{self.mmd_weight}")
        if self.method in ("fedadam", "fedyogi"):
            parts.append(f"server_lr_{self.server_lr}")
        if self.method in ("dynamic_layer", "sparse_coo"):
            parts.append(f"frac_{self.exchange_fraction}")
        return "_".join(parts)

    def hp_dir(self) -> Path:
        return Path(self.artifact_dir) / self.task / self.method / self.hp_key()


PENALISED = {"fedprox", "adaptive_fedprox", "ditto", "adaptive_ditto", "mr_mtl", "adaptive_mr_mtl", "moon", "perfcl", "fenda_ditto",
             "ditto_mkmmd", "mr_mtl_mkmmd", "ditto_deep_mmd", "mr_mtl_deep_mmd"}


def build(spec: ExperimentSpec, run_dir: Path, device: torch.device) -> tuple[Any, list[Any], Any]:
    from research.harness.methods import METHODS, MethodContext
    from research.harness.tasks import TASKS

    if spec.task not in TASKS:
        raise KeyError(f"unknown task '{spec.task}' (have: {sorted(TASKS)})")
    if spec.method not in METHODS:
        raise KeyError(f"unknown method '{spec.method}' (have: {sorted(METHODS)})")
    task = TASKS[spec.task](**(spec.task_kwargs or {}))
    ctx = MethodContext(spec, task, device, run_dir)
    server, clients = METHODS[spec.method](ctx)
    return server, clients, ctx


_SPMD_CONTEXT: Any = None


def _spmd_context() -> Any:
    """One process group per process, shared by every run of a sweep."""
    global _SPMD_CONTEXT
    if _SPMD_CONTEXT is None:
        from fl4health_b200.parallel.spmd import SpmdContext

        _SPMD_CONTEXT = SpmdContext()
    return _SPMD_CONTEXT


def _run_once(spec: ExperimentSpec, run_dir: Path, device: torch.device, spmd: bool) -> dict[str, Any]:
    from fl4health_b200.simulation import run_simulation

    run_dir.mkdir(parents=True, exist_ok=True)
    torch.manual_seed(spec.seed)
    rank = 0
    if spmd:
        from fl4health_b200.parallel.spmd import build_spmd_federation
        from fl4health_b200.parallel.spmd_multi import build_spmd_federation_multi

        ctx = _spmd_context()
        rank, device = ctx.rank, ctx.device
        server, clients, _ = build(spec, run_dir, device)
        assert len(clients) % ctx.world_size == 0, "the clients must divide evenly over the ranks"
        per_rank = len(clients) // ctx.world_size
        if rank != 0 and getattr(server.checkpoint_and_state_module, "model_checkpointers", None):
            server.checkpoint_and_state_module.model_checkpointers = None  # every rank holds the same global model: rank 0 saves it
        if per_rank == 1:
            build_spmd_federation(ctx, server, clients[rank])
        else:
            build_spmd_federation_multi(ctx, server, clients[rank * per_rank:(rank + 1) * per_rank])
        start = time.perf_counter()
        history, _ = server.fit(num_rounds=spec.rounds, timeout=None)
    else:
        server, clients, _ = build(spec, run_dir, device)
        start = time.perf_counter()
        history = run_simulation(server, clients, spec.rounds)
    elapsed = time.perf_counter() - start
    losses = [(int(r), float(v)) for r, v in history.losses_distributed]
    best = getattr(server, "best_aggregated_loss", None)
    if best is None and losses:
        best = min(v for _, v in losses)
    result = {"spec": asdict(spec), "losses_distributed": losses,
              "metrics_distributed": {k: [(int(r), float(v)) for r, v in vals] for k, vals in history.metrics_distributed.items()},
              "best_aggregated_loss": best, "seconds": elapsed, "rounds_per_s": spec.rounds / elapsed}
    if rank == 0:
        (run_dir / "results.json").write_text(json.dumps(result, indent=1))
        lines = [f"round {r}: weighted validation loss {v}" for r, v in losses]
        (run_dir / "server.out").write_text("\n".join(lines + [str(best)]) + "\n")
    return result


def run_experiment(spec: ExperimentSpec, device: torch.device | None = None, spmd: bool = False) -> list[dict[str, Any]]:
    """``spec.runs`` repetitions into ``<hp_dir>/Run{k}``; returns the per-run result dictionaries."""
    device = device or torch.device("cuda" if torch.cuda.is_available() else "cpu")
    results = []
    for run in range(1, spec.runs + 1):
        run_spec = replace(spec, seed=spec.seed + run - 1, data_seed=spec.data_seed + (run - 1 if spec.vary_data else 0))
        log(INFO, f"[research] {spec.task}/{spec.method}/{spec.hp_key()} run {run}/{spec.runs}")
        results.append(_run_once(run_spec, spec.hp_dir() / f"Run{run}", device, spmd))
    return results


def sweep(base: ExperimentSpec, grid: dict[str, list[Any]], device: torch.device | None = None, spmd: bool = False) -> dict[str, list[dict[str, Any]]]:
    """Cartesian product over ``grid`` (e.g. ``{"lr": [1e-3, 1e-2, 1e-1], "lam": [0.1, 1.0]}``)."""
    out = {}
    keys = sorted(grid)
    for values in itertools.product(*(grid[k] for k in keys)):
        spec = replace(base, **dict(zip(keys, values)))
        out[spec.hp_key()] = run_experiment(spec, device, spmd)
    return out
