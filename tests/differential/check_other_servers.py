"""The two one-shot servers vs the reference: model merging (clients send locally trained models once, the server averages
them, evaluates centrally and on the clients) and federated evaluation (clients score a checkpoint on their own data)."""
import importlib
import socket
import sys
import tempfile
import threading
from pathlib import Path

import torch
from torch import nn
from torch.utils.data import DataLoader

sys.path.insert(0, str(Path(__file__).resolve().parent))
import check_federations  # noqa: E402
from check_federations import BATCH, Net, cohort, pin_initialisation, resolver  # noqa: E402

SCRATCH = Path(tempfile.mkdtemp(prefix="fl4h_servers_"))


def local_model(index: int) -> nn.Module:
    """A 'pretrained' local model per site: the shared initialisation plus a site-specific perturbation."""
    model = pin_initialisation(Net())
    generator = torch.Generator().manual_seed(900 + index)
    with torch.no_grad():
        for parameter in model.parameters():
            parameter.add_(0.1 * torch.randn(parameter.shape, generator=generator))
    return model


def start(server, clients, ours: bool, rounds: int = 1):
    if ours:
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


def manager_for(side, ours: bool):
    return (side("servers.client_manager") if ours else importlib.import_module("flwr.server.client_manager")).SimpleClientManager()


def model_merge(prefix: str):
    side, ours = resolver(prefix), prefix == "fl4health_b200"
    dataset_module, accuracy = side("utils.dataset"), side("metrics").Accuracy
    client_cls = side("clients.model_merge_client").ModelMergeClient

    def hooks(index: int) -> dict:
        def get_model(self, config):
            return local_model(index).to(self.device)

        def get_test_data_loader(self, config):
            features, labels = cohort(index)
            return DataLoader(dataset_module.TensorDataset(features, labels), batch_size=BATCH, shuffle=False)

        return {"get_model": get_model, "get_test_data_loader": get_test_data_loader}

    clients = [type(f"Merge{i}", (client_cls,), hooks(i))(Path("."), SCRATCH / "unused.pt", [accuracy()], torch.device("cpu"), client_name=f"client_{i}") for i in range(3)]
    aggregation = side("metrics.metric_aggregation")
    features, labels = cohort(7)

    def central_evaluate(server_round, arrays, config):
        model = Net()
        tensors = [torch.as_tensor(a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a) for a in arrays]
        model.load_state_dict(dict(zip(model.state_dict(), tensors)))
        with torch.no_grad():
            logits = model(features)
        return float(nn.functional.cross_entropy(logits, labels)), {"central - accuracy": float((logits.argmax(1) == labels).float().mean())}

    config_fn = lambda server_round: {"current_server_round": server_round, "batch_size": BATCH}  # noqa: E731
    strategy = side("strategies.model_merge_strategy").ModelMergeStrategy(
        min_fit_clients=3, min_evaluate_clients=3, min_available_clients=3, on_fit_config_fn=config_fn, on_evaluate_config_fn=config_fn,
        fit_metrics_aggregation_fn=aggregation.fit_metrics_aggregation_fn, evaluate_metrics_aggregation_fn=aggregation.evaluate_metrics_aggregation_fn,
        evaluate_fn=central_evaluate)
    checkpointer = side("checkpointing.checkpointer").LatestTorchModuleCheckpointer(str(SCRATCH), f"merged_{prefix}.pt")
    server = side("servers.model_merge_server").ModelMergeServer(
        client_manager=manager_for(side, ours), strategy=strategy, checkpointer=checkpointer, server_model=Net(),
        parameter_exchanger=side("parameter_exchange.full_exchanger").FullParameterExchanger())
    history = start(server, clients, ours)
    merged = torch.load(SCRATCH / f"merged_{prefix}.pt", weights_only=False)
    return history, merged


def federated_evaluation(prefix: str):
    side, ours = resolver(prefix), prefix == "fl4health_b200"
    dataset_module, accuracy = side("utils.dataset"), side("metrics").Accuracy
    client_cls = side("clients.evaluate_client").EvaluateClient
    checkpoint = SCRATCH / "global_model.pt"
    torch.save(local_model(42), checkpoint)

    def hooks(index: int) -> dict:
        def initialize_global_model(self, config):
            return Net().to(self.device)  # architecture only: the weights arrive from the server

        def get_data_loader(self, config):
            features, labels = cohort(index)
            return (DataLoader(dataset_module.TensorDataset(features, labels), batch_size=BATCH, shuffle=False),)

        def get_criterion(self, config):
            return nn.CrossEntropyLoss()

        def get_local_model(self, config):
            return local_model(index).to(self.device)

        return dict(initialize_global_model=initialize_global_model, get_data_loader=get_data_loader, get_criterion=get_criterion, get_local_model=get_local_model)

    clients = [type(f"Eval{i}", (client_cls,), hooks(i))(Path("."), [accuracy()], torch.device("cpu"), client_name=f"client_{i}") for i in range(3)]
    manager = side("client_managers.fixed_without_replacement_manager").FixedSamplingByFractionClientManager()
    server = side("servers.evaluate_server").EvaluateServer(
        client_manager=manager, fraction_evaluate=1.0, model_checkpoint_path=checkpoint, evaluate_config={"batch_size": BATCH},
        evaluate_metrics_aggregation_fn=side("metrics.metric_aggregation").uniform_evaluate_metrics_aggregation_fn, accept_failures=False,
        min_available_clients=3)
    return start(server, clients, ours)


def series(history) -> dict:
    out = {"loss": history.losses_distributed, "loss_centralized": history.losses_centralized}
    for label in ("metrics_distributed_fit", "metrics_distributed", "metrics_centralized"):
        for key, values in getattr(history, label).items():
            out[f"{label}/{key}"] = values
    return out


def same_series(name: str, theirs, ours) -> None:
    a, b = series(theirs), series(ours)
    assert a.keys() == b.keys(), (name, sorted(a), sorted(b))
    for key in a:
        assert len(a[key]) == len(b[key]), (name, key, a[key], b[key])
        for (round_a, value_a), (round_b, value_b) in zip(a[key], b[key]):
            assert round_a == round_b and abs(float(value_a) - float(value_b)) < 2e-5, (name, key, a[key], b[key])
    assert any(len(values) > 0 for values in a.values()), (name, a)  # an empty history is not evidence
    print(f"  {name}: {sum(1 for v in a.values() if v)} non-empty series agree", file=sys.stderr)
    check_federations.agreed += 1


if __name__ == "__main__":
    (h_ref, merged_ref), (h_mine, merged_mine) = model_merge("fl4health"), model_merge("fl4health_b200")
    same_series("model_merge", h_ref, h_mine)
    for (name, a), (_, b) in zip(merged_ref.state_dict().items(), merged_mine.state_dict().items()):
        assert torch.allclose(a, b, atol=1e-6), name  # the checkpointed merged model
    same_series("federated_evaluation", federated_evaluation("fl4health"), federated_evaluation("fl4health_b200"))
    print("configs agree:", check_federations.agreed)
