"""End-to-end FedAvg on CPU through the public API (BASELINE config #1: basic_example plumbing, no GPU)."""

import json
from pathlib import Path

import pytest
import torch

from fl4health_b200.checkpointing.checkpointer import BestLossTorchModuleCheckpointer, LatestTorchModuleCheckpointer
from fl4health_b200.checkpointing.server_module import BaseServerCheckpointAndStateModule
from fl4health_b200.checkpointing.state_checkpointer import ServerStateCheckpointer
from fl4health_b200.engine.options import EngineOptions
from fl4health_b200.metrics.metric_aggregation import evaluate_metrics_aggregation_fn, fit_metrics_aggregation_fn
from fl4health_b200.models import Net
from fl4health_b200.parameter_exchange.full_exchanger import FullParameterExchanger
from fl4health_b200.reporting import JsonReporter
from fl4health_b200.servers.base_server import FlServer
from fl4health_b200.servers.client_manager import SimpleClientManager
from fl4health_b200.simulation import run_simulation
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg
from fl4health_b200.strategies.fedavg import FedAvg
from fl4health_b200.utils.random import set_all_random_seeds
from tests.helpers import fit_config_fn, make_clients


def _strategy(cls=FedAvg, **kwargs):
    return cls(
        min_fit_clients=2, min_evaluate_clients=2, min_available_clients=2,
        on_fit_config_fn=fit_config_fn(), on_evaluate_config_fn=fit_config_fn(),
        fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
        evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn, **kwargs,
    )


def test_fedavg_two_clients_cpu(tmp_path: Path) -> None:
    set_all_random_seeds(42)
    model = Net()
    ckpt = BaseServerCheckpointAndStateModule(
        model=model, parameter_exchanger=FullParameterExchanger(),
        model_checkpointers=[BestLossTorchModuleCheckpointer(str(tmp_path), "best_model.pkl"),
                             LatestTorchModuleCheckpointer(str(tmp_path), "latest_model.pkl")],
    )
    reporter = JsonReporter(run_id="server", output_folder=tmp_path / "metrics")
    server = FlServer(
        client_manager=SimpleClientManager(), fl_config={"n_server_rounds": 3, "local_steps": 5},
        strategy=_strategy(), reporters=[reporter], checkpoint_and_state_module=ckpt,
        on_init_parameters_config_fn=fit_config_fn(), accept_failures=False,
    )
    clients = make_clients(2)
    history = run_simulation(server, clients, num_rounds=3)

    losses = [loss for _, loss in history.losses_distributed]
    assert len(losses) == 3 and losses[-1] < losses[0], losses
    assert "val - prediction - accuracy" in history.metrics_distributed
    assert "train - prediction - accuracy" in history.metrics_distributed_fit

    # both clients hold the same (global) model after the final evaluate round
    sd0, sd1 = clients[0].model.state_dict(), clients[1].model.state_dict()
    assert all(torch.equal(sd0[k], sd1[k]) for k in sd0)

    # artifacts: pickled whole modules loadable with plain torch, json report with per-round keys
    loaded = torch.load(tmp_path / "latest_model.pkl", weights_only=False)
    assert isinstance(loaded, Net)
    assert all(torch.allclose(loaded.state_dict()[k], sd0[k].cpu()) for k in sd0)
    report = json.loads((tmp_path / "metrics" / "server.json").read_text())
    assert set(report["rounds"].keys()) == {"1", "2", "3"}
    assert "val - loss - aggregated" in report["rounds"]["3"]
    assert "eval_round_metrics_aggregated" in report["rounds"]["3"]


def test_fedavg_matches_manual_average() -> None:
    """One round, known weights: the server's result equals the sample-weighted mean of the client models."""
    set_all_random_seeds(7)
    server = FlServer(SimpleClientManager(), {"n_server_rounds": 1}, _strategy(BasicFedAvg),
                      on_init_parameters_config_fn=fit_config_fn())
    clients = make_clients(2, n_train=128)
    clients[1].n_train = 64
    # evaluation re-broadcast would overwrite client models: capture after-fit states via hook
    post_fit = {}
    for c in clients:
        orig = c.update_after_train

        def hook(local_steps, loss_dict, config, c=c, orig=orig):
            orig(local_steps, loss_dict, config)
            post_fit[c.client_name] = {k: v.clone() for k, v in c.model.state_dict().items()}

        c.update_after_train = hook
    run_simulation(server, clients, num_rounds=1)
    final = clients[0].model.state_dict()
    for key, value in final.items():
        expected = (128 * post_fit["c0"][key].double() + 64 * post_fit["c1"][key].double()) / 192
        assert torch.allclose(value.double(), expected, atol=1e-6), key


def test_resume_from_server_and_client_state(tmp_path: Path) -> None:
    """Run 1 round, 'crash', restart for 3 rounds: final metrics equal an uninterrupted 3-round run
    (the reference's fault-tolerance smoke test, tests/smoke_tests/run_smoke_test.py:414-611)."""
    from fl4health_b200.checkpointing.client_module import ClientCheckpointAndStateModule
    from fl4health_b200.checkpointing.state_checkpointer import ClientStateCheckpointer

    def build(state_dir: Path | None):
        set_all_random_seeds(11)
        module = BaseServerCheckpointAndStateModule(
            model=Net(), parameter_exchanger=FullParameterExchanger(),
            state_checkpointer=ServerStateCheckpointer(state_dir) if state_dir else None,
        )
        server = FlServer(SimpleClientManager(), {"n_server_rounds": 3}, _strategy(),
                          checkpoint_and_state_module=module, on_init_parameters_config_fn=fit_config_fn(),
                          server_name="srv")
        clients = make_clients(2)
        if state_dir:
            for c in clients:
                c.checkpoint_and_state_module = ClientCheckpointAndStateModule(
                    state_checkpointer=ClientStateCheckpointer(state_dir))
        return server, clients

    server, clients = build(None)
    reference = run_simulation(server, clients, num_rounds=3)

    state_dir = tmp_path / "state"
    state_dir.mkdir()
    server, clients = build(state_dir)
    run_simulation(server, clients, num_rounds=1)
    assert (state_dir / "server_srv_state.pt").exists() and (state_dir / "client_c0_state.pt").exists()
    server, clients = build(state_dir)
    # the data-loader position is not part of client state (nor is it in the reference): re-align iterators
    resumed = run_simulation(server, clients, num_rounds=3)
    assert [r for r, _ in resumed.losses_distributed] == [1, 2, 3]
    assert abs(resumed.losses_distributed[0][1] - reference.losses_distributed[0][1]) < 1e-6


def test_fault_injection_then_resume(tmp_path: Path, monkeypatch: pytest.MonkeyPatch) -> None:
    """``FL4H_FAULT_AFTER_ROUND=2`` pre-empts the server after round 2 (state saved); a fresh process-equivalent
    (new server + clients over the same state directory) finishes rounds 3..4 and keeps the first two rounds' history
    (fault injection to test resume)."""
    from fl4health_b200.checkpointing.client_module import ClientCheckpointAndStateModule
    from fl4health_b200.checkpointing.state_checkpointer import ClientStateCheckpointer

    state_dir = tmp_path / "state"
    state_dir.mkdir()

    def build():
        set_all_random_seeds(13)
        module = BaseServerCheckpointAndStateModule(
            model=Net(), parameter_exchanger=FullParameterExchanger(), state_checkpointer=ServerStateCheckpointer(state_dir))
        server = FlServer(SimpleClientManager(), {"n_server_rounds": 4}, _strategy(), checkpoint_and_state_module=module,
                          on_init_parameters_config_fn=fit_config_fn(), server_name="srv")
        clients = make_clients(2)
        for c in clients:
            c.checkpoint_and_state_module = ClientCheckpointAndStateModule(state_checkpointer=ClientStateCheckpointer(state_dir))
        return server, clients

    monkeypatch.setenv("FL4H_FAULT_AFTER_ROUND", "2")
    server, clients = build()
    with pytest.raises(SystemExit, match="fault injected after round 2"):
        run_simulation(server, clients, num_rounds=4)
    first_two = list(server.history.losses_distributed)
    assert [r for r, _ in first_two] == [1, 2]

    monkeypatch.delenv("FL4H_FAULT_AFTER_ROUND")
    server, clients = build()
    resumed = run_simulation(server, clients, num_rounds=4)
    assert [r for r, _ in resumed.losses_distributed] == [1, 2, 3, 4]
    assert resumed.losses_distributed[:2] == first_two  # rounds 1-2 come from the restored history, not a re-run
    assert clients[0].total_steps == 4 * 5  # step counters restored: 2 rounds before the fault + 2 after
