"""Checkpoint / resume vs the reference: a federation that stops after 2 of 4 rounds and is restarted with fresh server
and client objects pointed at the same state directory.  Both implementations must (a) leave the same set of files,
(b) continue at round 3, (c) end with the same 4-round history -- which, on our side, is also the history of an
uninterrupted run."""
import importlib
import socket
import sys
import tempfile
import threading
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
import check_federations  # noqa: E402
from check_federations import CLIENTS, Net, compare, initial_parameters, pin_initialisation, resolver, round_config, user_hooks  # noqa: E402


def build(prefix: str, state_dir: Path | None):
    side, ours = resolver(prefix), prefix == "fl4health_b200"
    accuracy = side("metrics").Accuracy
    client_cls = side("clients.basic_client").BasicClient
    clients = []
    for index in range(CLIENTS):
        module = None
        if state_dir is not None:
            checkpointing = side("checkpointing.checkpointer")
            module = side("checkpointing.client_module").ClientCheckpointAndStateModule(
                pre_aggregation=[checkpointing.LatestTorchModuleCheckpointer(str(state_dir), f"client_{index}_pre_latest.pkl")],
                post_aggregation=[checkpointing.BestLossTorchModuleCheckpointer(str(state_dir), f"client_{index}_post_best.pkl")],
                state_checkpointer=side("checkpointing.state_checkpointer").ClientStateCheckpointer(state_dir))
        cls = type(f"Client{index}", (client_cls,), user_hooks(side, index))
        clients.append(cls(data_path=Path("."), metrics=[accuracy()], device=torch.device("cpu"), client_name=f"client_{index}",
                           checkpoint_and_state_module=module))
    aggregation = side("metrics.metric_aggregation")
    config_fn = round_config({})
    strategy = side("strategies.basic_fedavg").BasicFedAvg(
        min_fit_clients=CLIENTS, min_evaluate_clients=CLIENTS, min_available_clients=CLIENTS, on_fit_config_fn=config_fn,
        on_evaluate_config_fn=config_fn, fit_metrics_aggregation_fn=aggregation.fit_metrics_aggregation_fn,
        evaluate_metrics_aggregation_fn=aggregation.evaluate_metrics_aggregation_fn, **initial_parameters()(side, ours))
    server_module = None
    if state_dir is not None:
        checkpointing = side("checkpointing.checkpointer")
        server_module = side("checkpointing.server_module").BaseServerCheckpointAndStateModule(
            model=pin_initialisation(Net()), parameter_exchanger=side("parameter_exchange.full_exchanger").FullParameterExchanger(),
            model_checkpointers=[checkpointing.BestLossTorchModuleCheckpointer(str(state_dir), "best_model.pkl"),
                                 checkpointing.LatestTorchModuleCheckpointer(str(state_dir), "latest_model.pkl")],
            state_checkpointer=side("checkpointing.state_checkpointer").ServerStateCheckpointer(state_dir))
    manager = (side("servers.client_manager") if ours else importlib.import_module("flwr.server.client_manager")).SimpleClientManager()
    server = side("servers.base_server").FlServer(client_manager=manager, fl_config={"n_server_rounds": 4}, strategy=strategy,
                                                  checkpoint_and_state_module=server_module, server_name="server", accept_failures=False)
    return server, clients


def run(prefix: str, state_dir: Path | None, rounds: int):
    server, clients = build(prefix, state_dir)
    if prefix == "fl4health_b200":
        from fl4health_b200.simulation import run_simulation

        return run_simulation(server, clients, num_rounds=rounds)
    import flwr

    with socket.socket() as probe:
        probe.bind(("127.0.0.1", 0))
        address = f"127.0.0.1:{probe.getsockname()[1]}"
    threads = [threading.Thread(target=flwr.client.start_client, kwargs=dict(server_address=address, client=c.to_client(), cid=c.client_name), daemon=True) for c in clients]
    for thread in threads:
        thread.start()
    history = flwr.server.start_server(server=server, server_address=address, config=flwr.server.ServerConfig(num_rounds=rounds))
    for thread in threads:
        thread.join(60)
    return history


if __name__ == "__main__":
    outcomes, files, models = {}, {}, {}
    for prefix in ("fl4health", "fl4health_b200"):
        state_dir = Path(tempfile.mkdtemp(prefix=f"state_{prefix}_"))
        first = run(prefix, state_dir, rounds=2)
        assert [r for r, _ in first.losses_distributed] == [1, 2]
        files[prefix] = sorted(p.name for p in state_dir.iterdir())
        resumed = run(prefix, state_dir, rounds=4)  # new objects, same directory
        assert [r for r, _ in resumed.losses_distributed] == [1, 2, 3, 4], (prefix, resumed.losses_distributed)
        outcomes[prefix] = resumed
        models[prefix] = {p.name: torch.load(p, weights_only=False).state_dict() for p in sorted(state_dir.glob("*.pkl")) if "model" in p.name or "client_" in p.name}
    assert files["fl4health"] == files["fl4health_b200"], files  # same artefacts under the same names
    compare("resume after 2 of 4 rounds", outcomes["fl4health"], outcomes["fl4health_b200"], tol=2e-4)
    # the model checkpoints both sides left behind (server: best / latest; clients: latest before, best after aggregation)
    assert models["fl4health"].keys() == models["fl4health_b200"].keys() and len(models["fl4health"]) == 2 + 2 * CLIENTS, sorted(models["fl4health"])
    for name, state in models["fl4health"].items():
        for (key, a), (_, b) in zip(state.items(), models["fl4health_b200"][name].items()):
            assert torch.allclose(a, b, atol=2e-4), (name, key, (a - b).abs().max())
    check_federations.agreed += 1
    uninterrupted = run("fl4health_b200", Path(tempfile.mkdtemp(prefix="state_uninterrupted_")), rounds=4)  # same modules, no stop
    compare("resumed == uninterrupted (ours)", uninterrupted, outcomes["fl4health_b200"], tol=1e-6)
    print("configs agree:", check_federations.agreed)
