"""A tabular federation with feature alignment, end to end, vs the reference: hospitals whose frames differ (missing
columns, categories unseen elsewhere) negotiate a schema with the server (given a source of truth, or polled from a
client), align, and train a model sized from the aligned dimensions.  Compared: aligned dimensions, the schema the server
holds, per-round histories."""
import json
import socket
import sys
import threading
from pathlib import Path

import numpy as np
import pandas as pd
import torch
from torch import nn
from torch.utils.data import DataLoader

sys.path.insert(0, str(Path(__file__).resolve().parent))
from check_federations import compare, pin_initialisation, resolver  # noqa: E402

import check_federations  # noqa: E402

ROUNDS, BATCH = 3, 16


def hospital_frame(index: int, n: int = 120) -> pd.DataFrame:
    rng = np.random.default_rng(500 + index)
    wards = ["icu", "er", "surgery", "medicine"] + (["oncology"] if index == 2 else [])
    frame = pd.DataFrame({
        "patient": np.arange(n) + 1000 * index,
        "age": rng.integers(18, 90, n).astype(float),
        "lactate": rng.normal(2.0 + 0.3 * index, 0.7, n),
        "sex": rng.choice(["F", "M"], n),
        "ward": rng.choice(wards, n),
        "triage": rng.integers(1, 6, n),
        "note": rng.choice(["stable overnight", "fever and cough", "chest pain on arrival", "post operative day two"], n),
    })
    frame["outcome"] = ((frame["lactate"] > 2.2).astype(int) + (frame["age"] > 60).astype(int)).astype(int)
    frame.loc[rng.choice(n, n // 12, replace=False), "lactate"] = np.nan
    if index == 1:
        frame = frame.drop(columns=["triage"])  # this site never recorded it
    return frame


class Mlp(nn.Module):
    def __init__(self, inputs: int, outputs: int) -> None:
        super().__init__()
        self.layers = nn.Sequential(nn.Linear(inputs, 12), nn.ReLU(), nn.Linear(12, outputs))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.layers(x)


def build(side, ours: bool, source_specified: bool):
    client_cls = side("clients.tabular_data_client").TabularDataClient
    dataset_module = side("utils.dataset")
    accuracy = side("metrics").Accuracy
    dimensions = {}

    def hooks(index: int) -> dict:
        def get_data_frame(self, config):
            return hospital_frame(index)

        def get_data_loaders(self, config):
            features = self.aligned_features.toarray() if hasattr(self.aligned_features, "toarray") else np.asarray(self.aligned_features)
            x = torch.from_numpy(np.asarray(features, dtype=np.float32))
            x = torch.nan_to_num((x - x.mean(dim=0)) / (x.std(dim=0) + 1.0))
            y = torch.from_numpy(np.asarray(self.aligned_targets)).long().reshape(-1)
            dimensions[index] = (self.input_dimension, self.output_dimension)
            return (DataLoader(dataset_module.TensorDataset(x[:96], y[:96]), batch_size=BATCH, shuffle=False),
                    DataLoader(dataset_module.TensorDataset(x[96:], y[96:]), batch_size=BATCH, shuffle=False))

        def get_model(self, config):
            return pin_initialisation(Mlp(self.input_dimension, self.output_dimension)).to(self.device)

        def get_optimizer(self, config):
            return torch.optim.SGD(self.model.parameters(), lr=0.05, momentum=0.9)

        def get_criterion(self, config):
            return nn.CrossEntropyLoss()

        return dict(get_data_frame=get_data_frame, get_data_loaders=get_data_loaders, get_model=get_model, get_optimizer=get_optimizer, get_criterion=get_criterion)

    clients = [type(f"Hospital{i}", (client_cls,), hooks(i))(Path("."), [accuracy()], torch.device("cpu"), "patient", "outcome", client_name=f"hospital_{i}")
               for i in range(3)]
    aggregation = side("metrics.metric_aggregation")
    strategy = side("strategies.basic_fedavg").BasicFedAvg(
        min_fit_clients=3, min_evaluate_clients=3, min_available_clients=3, on_fit_config_fn=None, on_evaluate_config_fn=None,
        fit_metrics_aggregation_fn=aggregation.fit_metrics_aggregation_fn, evaluate_metrics_aggregation_fn=aggregation.evaluate_metrics_aggregation_fn,
        initial_parameters=None)
    encoder_cls = side("feature_alignment.tab_features_info_encoder").TabularFeaturesInfoEncoder
    source = encoder_cls.encoder_from_dataframe(hospital_frame(0), "patient", "outcome") if source_specified else None
    extraction = side("utils.parameter_extraction")
    # (a fraction-sampling manager, as in the reference's example: its poll of a plain Flower manager sizes the sample before
    # the clients have connected)
    manager = side("client_managers.fixed_without_replacement_manager").FixedSamplingByFractionClientManager()
    server = side("servers.tabular_feature_alignment_server").TabularFeatureAlignmentServer(
        client_manager=manager, config={"n_server_rounds": ROUNDS, "local_steps": 4, "batch_size": BATCH},
        initialize_parameters=lambda inputs, outputs: extraction.get_all_model_parameters(pin_initialisation(Mlp(inputs, outputs))),
        strategy=strategy, tabular_features_source_of_truth=source, accept_failures=False)
    return server, clients, dimensions


def run(prefix: str, source_specified: bool):
    ours = prefix == "fl4health_b200"
    server, clients, dimensions = build(resolver(prefix), ours, source_specified)
    if ours:
        from fl4health_b200.simulation import run_simulation

        history = run_simulation(server, clients, num_rounds=ROUNDS)
    else:
        import flwr

        with socket.socket() as probe:
            probe.bind(("127.0.0.1", 0))
            address = f"127.0.0.1:{probe.getsockname()[1]}"
        threads = [threading.Thread(target=flwr.client.start_client, kwargs=dict(server_address=address, client=c.to_client(), cid=c.client_name), daemon=True) for c in clients]
        for thread in threads:
            thread.start()
        history = flwr.server.start_server(server=server, server_address=address, config=flwr.server.ServerConfig(num_rounds=ROUNDS))
        for thread in threads:
            thread.join(60)
    schema = server.tab_features_info.to_json() if server.tab_features_info is not None else server.fl_config.get("feature_info")
    return history, dimensions, dict(server.dimension_info), schema


if __name__ == "__main__":
    for source_specified in (True,):
        h_ref, d_ref, info_ref, schema_ref = run("fl4health", source_specified)
        h_mine, d_mine, info_mine, schema_mine = run("fl4health_b200", source_specified)
        assert d_ref == d_mine and len(set(d_ref.values())) == 1, (d_ref, d_mine)  # every site lands in the same space
        assert info_ref == info_mine, (info_ref, info_mine)
        assert json.loads(schema_ref) == json.loads(schema_mine)
        compare(f"tabular_alignment(source_specified={source_specified})", h_ref, h_mine, tol=2e-4)
    print("configs agree:", check_federations.agreed)
