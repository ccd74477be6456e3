"""Model selection and held-out evaluation over the artefacts written by ``research.harness.experiment``.

* ``find_best_hp`` — the hyper-parameter folder whose runs have the lowest mean final-line loss in ``server.out``
  (parity: ``research/cifar10/find_best_hp.py:9-44``).
* ``evaluate_on_test`` — loads every run's client (and, when present, server) checkpoints, evaluates them on each
  client's held-out test split and reports mean ± std over runs (parity: ``research/cifar10/evaluate_on_test.py``,
  ``research/flamby/utils.py``).
"""

from __future__ import annotations

import json
import math
from logging import INFO
from pathlib import Path
from typing import Any

import torch
from torch import nn

from fl4health_b200.common.logger import log


def get_hp_folders(hp_sweep_dir: str | Path) -> list[Path]:
    return sorted(p for p in Path(hp_sweep_dir).iterdir() if p.is_dir())


def get_run_folders(hp_dir: str | Path) -> list[Path]:
    return sorted(p for p in Path(hp_dir).iterdir() if p.is_dir() and "Run" in p.name)


def get_weighted_loss_from_server_log(run_folder: str | Path) -> float:
    lines = [line.strip() for line in (Path(run_folder) / "server.out").read_text().splitlines() if line.strip()]
    return float(lines[-1])


def find_best_hp(hp_sweep_dir: str | Path) -> tuple[Path, float]:
    best_dir, best_loss = None, math.inf
    for hp_folder in get_hp_folders(hp_sweep_dir):
        runs = get_run_folders(hp_folder)
        if not runs:
            continue
        mean_loss = sum(get_weighted_loss_from_server_log(r) for r in runs) / len(runs)
        log(INFO, f"{hp_folder.name}: mean loss over {len(runs)} run(s) = {mean_loss}")
        if mean_loss <= best_loss:
            best_dir, best_loss = hp_folder, mean_loss
    if best_dir is None:
        raise FileNotFoundError(f"no hyper-parameter folders with runs under {hp_sweep_dir}")
    log(INFO, f"Best Loss: {best_loss}\nBest Folder: {best_dir}")
    return best_dir, best_loss


def _prediction(output: Any) -> torch.Tensor:
    if isinstance(output, tuple):
        output = output[0]
    if isinstance(output, dict):
        for key in ("prediction", "personal", "local", "global"):
            if key in output:
                return output[key]
        return next(iter(output.values()))
    return output


@torch.no_grad()
def evaluate_model(model: nn.Module, dataset: Any, device: torch.device, batch_size: int = 256) -> dict[str, float]:
    model = model.to(device).eval()
    correct, loss_sum, n = 0, 0.0, len(dataset.data)
    for start in range(0, n, batch_size):
        x, y = dataset.data[start:start + batch_size].to(device), dataset.targets[start:start + batch_size].to(device)
        logits = _prediction(model(x)).float()
        loss_sum += float(torch.nn.functional.cross_entropy(logits, y, reduction="sum"))
        correct += int((logits.argmax(dim=1) == y).sum())
    return {"accuracy": correct / max(n, 1), "loss": loss_sum / max(n, 1), "n": n}


def _mean_std(values: list[float]) -> tuple[float, float]:
    mean = sum(values) / len(values)
    return mean, math.sqrt(sum((v - mean) ** 2 for v in values) / len(values))


def evaluate_on_test(hp_dir: str | Path, device: torch.device | None = None, which: str = "best") -> dict[str, Any]:
    """Test-set performance of the checkpoints under ``hp_dir/Run*``: per client, client-averaged, mean ± std over
    runs; for full-exchange methods also the server model's client-averaged accuracy."""
    from research.harness.experiment import ExperimentSpec
    from research.harness.tasks import TASKS

    device = device or torch.device("cuda" if torch.cuda.is_available() else "cpu")
    per_run_client, per_run_server, per_client = [], [], {}
    for run in get_run_folders(hp_dir):
        spec = ExperimentSpec(**json.loads((run / "results.json").read_text())["spec"])
        task = TASKS[spec.task](**(spec.task_kwargs or {}))
        tests = [task.client_data(i, spec)[2] for i in range(task.n_clients)]
        accuracies = []
        for index, test in enumerate(tests):
            path = run / f"client_{index}_{which}_model.pkl"
            if not path.exists():  # e.g. `central` has a single participant
                continue
            outcome = evaluate_model(torch.load(path, weights_only=False, map_location="cpu"), test, device)
            accuracies.append(outcome["accuracy"])
            per_client.setdefault(f"client_{index}", []).append(outcome["accuracy"])
        if accuracies:
            per_run_client.append(sum(accuracies) / len(accuracies))
        server_path = run / f"server_{which}_model.pkl"
        if server_path.exists():
            model = torch.load(server_path, weights_only=False, map_location="cpu")
            per_run_server.append(sum(evaluate_model(model, t, device)["accuracy"] for t in tests) / len(tests))
    report: dict[str, Any] = {"hp_dir": str(hp_dir), "runs": len(get_run_folders(hp_dir)), "checkpoint": which}
    if per_run_client:
        report["client_models_avg_accuracy"], report["client_models_std"] = _mean_std(per_run_client)
        report["per_client_accuracy"] = {k: _mean_std(v)[0] for k, v in sorted(per_client.items())}
    if per_run_server:
        report["server_model_avg_accuracy"], report["server_model_std"] = _mean_std(per_run_server)
    (Path(hp_dir) / f"test_eval_{which}.json").write_text(json.dumps(report, indent=1))
    return report
