"""What a federation reports: the JSON documents written by client and server reporters over a two-round run
(modelled on the reporter-payload checks of the reference's ``tests/clients/test_basic_client.py`` and
``tests/servers/test_base_server.py:153-203``)."""

from __future__ import annotations

import json
from pathlib import Path

import torch

from fl4health_b200.metrics import Accuracy
from fl4health_b200.metrics.metric_aggregation import evaluate_metrics_aggregation_fn, fit_metrics_aggregation_fn
from fl4health_b200.reporting import JsonReporter
from fl4health_b200.servers.base_server import FlServer
from fl4health_b200.servers.client_manager import SimpleClientManager
from fl4health_b200.simulation import run_simulation
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg
from tests.helpers import SyntheticCifarClient, TinyNet


def _config(server_round: int) -> dict:
    return {"current_server_round": server_round, "local_steps": 3, "batch_size": 16}


def _federation(tmp_path: Path, n_clients: int = 2, **fit_extra):  # type: ignore[no-untyped-def]
    def fn(server_round: int) -> dict:
        return {**_config(server_round), **fit_extra}

    clients = [SyntheticCifarClient(Path("."), [Accuracy()], torch.device("cpu"), client_name=f"c{i}", seed=i, n_train=64, n_val=32,
                                    model_fn=TinyNet, reporters=[JsonReporter(run_id=f"client_{i}", output_folder=tmp_path)])
               for i in range(n_clients)]
    strategy = BasicFedAvg(min_fit_clients=n_clients, min_evaluate_clients=n_clients, min_available_clients=n_clients, on_fit_config_fn=fn,
                           on_evaluate_config_fn=fn, fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
                           evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn)
    server = FlServer(SimpleClientManager(), {"n_server_rounds": 2}, strategy, reporters=[JsonReporter(run_id="server", output_folder=tmp_path)],
                      on_init_parameters_config_fn=fn)
    return server, clients


def test_client_and_server_json_reports(tmp_path: Path) -> None:
    server, clients = _federation(tmp_path)
    run_simulation(server, clients, 2)

    report = json.loads((tmp_path / "client_0.json").read_text())
    assert report["host_type"] == "client" and "initialized" in report and "shutdown" in report
    rounds = report["rounds"]
    assert sorted(rounds) == ["1", "2"]
    for index, (key, payload) in enumerate(sorted(rounds.items()), start=1):
        assert payload["round"] == index
        for field in ("round_start", "round_end", "fit_round_start", "fit_round_end", "fit_round_time_elapsed", "eval_round_start",
                      "eval_round_end", "eval_round_time_elapsed"):
            assert field in payload, f"round {key}: missing {field}"
        assert payload["fit_step"] == 3 * index            # cumulative optimisation steps
        assert set(payload["fit_round_losses"]) == {"backward"} and payload["fit_round_losses"]["backward"] > 0
        assert "train - prediction - accuracy" in payload["fit_round_metrics"]
        assert "val - prediction - accuracy" in payload["eval_round_metrics"] and payload["eval_round_loss"] > 0
        assert payload["round_start"] <= payload["fit_round_start"] <= payload["fit_round_end"] <= payload["round_end"]

    server_report = json.loads((tmp_path / "server.json").read_text())
    assert server_report["host_type"] == "server" and server_report["num_rounds"] == 2
    for field in ("fit_start", "fit_end", "fit_elapsed_time", "shutdown"):
        assert field in server_report
    for key in ("1", "2"):
        payload = server_report["rounds"][key]
        assert payload["fit_round_start"] <= payload["fit_round_end"]
        assert "train - prediction - accuracy" in payload["fit_round_metrics"]
        assert "val - prediction - accuracy" in payload["eval_round_metrics_aggregated"]


def test_step_level_reporting_and_evaluate_after_fit(tmp_path: Path) -> None:
    server, clients = _federation(tmp_path, n_clients=1, evaluate_after_fit=True)
    clients[0].reports_manager.reporters[0].run_id = "stepwise"
    run_simulation(server, clients, 2)
    report = json.loads((tmp_path / "stepwise.json").read_text())
    first = report["rounds"]["1"]
    # evaluate_after_fit folds validation metrics of the locally trained model into the fit results
    assert any(name.startswith("val - ") for name in first["fit_round_metrics"])
    # per-step payloads are nested under the round (reference: rounds -> steps -> fit_step_losses)
    steps = first.get("steps", {})
    if steps:
        assert sorted(int(s) for s in steps) == [1, 2, 3] and all("fit_step_losses" in payload for payload in steps.values())
