"""Tabular feature alignment: schema inference / JSON round trip, aligned shapes across heterogeneous clients, and
the two-poll server protocol end to end."""

from pathlib import Path

import numpy as np
import pandas as pd
import pytest
import torch
from torch import nn

from fl4health_b200.clients.tabular_data_client import TabularDataClient
from fl4health_b200.common.typing import ndarrays_to_parameters
from fl4health_b200.engine.data import BatchedTensorLoader
from fl4health_b200.feature_alignment.constants import FeatureType
from fl4health_b200.feature_alignment.handle_types import infer_types, to_types
from fl4health_b200.feature_alignment.tab_features_info_encoder import TabularFeaturesInfoEncoder
from fl4health_b200.feature_alignment.tab_features_preprocessor import TabularFeaturesPreprocessor
from fl4health_b200.feature_alignment.tabular_feature import TabularFeature
from fl4health_b200.feature_alignment.tabular_type import TabularType
from fl4health_b200.metrics import Accuracy
from fl4health_b200.metrics.metric_aggregation import evaluate_metrics_aggregation_fn, fit_metrics_aggregation_fn
from fl4health_b200.servers.client_manager import SimpleClientManager
from fl4health_b200.servers.tabular_feature_alignment_server import TabularFeatureAlignmentServer
from fl4health_b200.simulation import run_simulation
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg
from fl4health_b200.utils.dataset import TensorDataset
from fl4health_b200.utils.random import set_all_random_seeds


def _frame(seed: int, n: int = 120, drop: str | None = None) -> pd.DataFrame:
    rng = np.random.default_rng(seed)
    age = rng.normal(50, 10, n)
    smoker = rng.integers(0, 2, n)
    df = pd.DataFrame({
        "pid": np.arange(n) + 1000 * seed,
        "age": age,
        "smoker": smoker.astype(bool),
        "ward": rng.choice(["icu", "er", "general"], n),
        "note": rng.choice(["patient stable", "needs oxygen", "stable after oxygen", "critical condition"], n),
        "outcome": ((age > 50) ^ (smoker == 1)).astype(int),
    })
    return df.drop(columns=[drop]) if drop else df


def test_type_inference_and_schema_round_trip() -> None:
    df = _frame(0)
    types = infer_types(df, ["age", "smoker", "ward", "note", "outcome"])
    assert types == {"age": FeatureType.NUMERIC, "smoker": FeatureType.BINARY, "ward": FeatureType.ORDINAL,
                     "note": FeatureType.ORDINAL, "outcome": FeatureType.BINARY}
    wide = pd.DataFrame({"txt": [f"free text number {i}" for i in range(30)], "x": np.arange(30) * 0.5})
    assert infer_types(wide, ["txt", "x"]) == {"txt": FeatureType.STRING, "x": FeatureType.NUMERIC}
    converted, meta = to_types(df, {"ward": FeatureType.ORDINAL, "age": FeatureType.NUMERIC})
    assert set(converted["ward"].unique()) == {0, 1, 2} and meta["ward"]["mapping"][0] == "er"
    encoder = TabularFeaturesInfoEncoder.encoder_from_dataframe(df, "pid", "outcome")
    assert encoder.get_feature_columns() == ["age", "note", "smoker", "ward"] and encoder.get_target_columns() == ["outcome"]
    assert encoder.get_target_dimension() == 2
    again = TabularFeaturesInfoEncoder.from_json(encoder.to_json())
    assert [f.get_metadata() for f in again.get_tabular_features()] == [f.get_metadata() for f in encoder.get_tabular_features()]
    feature = TabularFeature("ward", TabularType.ORDINAL, None, ["er", "general", "icu"])
    assert TabularFeature.from_json(feature.to_json()).get_fill_value() == "UNKNOWN"
    assert TabularType.get_default_fill_value(TabularType.NUMERIC) == 0.0
    with pytest.raises(ValueError):
        TabularFeature("t", TabularType.STRING, None, {"a": 0}).get_metadata_dimension()


def test_alignment_gives_identical_shapes_even_with_missing_columns() -> None:
    schema = TabularFeaturesInfoEncoder.encoder_from_dataframe(_frame(0), "pid", "outcome")
    x0, y0 = TabularFeaturesPreprocessor(schema).preprocess_features(_frame(0))
    x1, y1 = TabularFeaturesPreprocessor(schema).preprocess_features(_frame(1, n=80, drop="ward"))  # client lacks a column
    dense = lambda a: a.toarray() if hasattr(a, "toarray") else a  # noqa: E731
    assert dense(x0).shape[1] == dense(x1).shape[1] == 1 + 4 + 1 + 3  # age + one-hot(note) + smoker + one-hot(ward)
    assert dense(x1)[:, -3:].sum() == 0  # unknown ward -> all-zero one-hot block
    assert y0.shape == (120, 1) and y1.shape == (80, 1)


class _TabClient(TabularDataClient):
    def __init__(self, seed: int, drop: str | None, **kwargs) -> None:
        super().__init__(Path("."), [Accuracy()], torch.device("cpu"), id_column="pid", targets="outcome", **kwargs)
        self.seed, self.drop = seed, drop

    def get_data_frame(self, config):
        return _frame(self.seed, drop=self.drop)

    def get_data_loaders(self, config):
        x = torch.from_numpy(np.asarray(self.aligned_features, dtype=np.float32))
        y = torch.from_numpy(np.asarray(self.aligned_targets)).long().reshape(-1)
        split = int(0.8 * len(x))
        bs = int(config["batch_size"])
        return (BatchedTensorLoader(TensorDataset(x[:split], y[:split]), bs, shuffle=True),
                BatchedTensorLoader(TensorDataset(x[split:], y[split:]), bs))

    def get_model(self, config):
        return nn.Sequential(nn.Linear(self.input_dimension, 16), nn.ReLU(), nn.Linear(16, self.output_dimension))

    def get_criterion(self, config):
        return nn.CrossEntropyLoss()

    def get_optimizer(self, config):
        return torch.optim.SGD(self.model.parameters(), lr=0.1)


@pytest.mark.parametrize("server_has_schema", [False, True])
def test_tabular_alignment_federation(server_has_schema: bool) -> None:
    set_all_random_seeds(5)
    clients = [_TabClient(1, None, client_name="t0"), _TabClient(2, "ward", client_name="t1")]
    dims = {}

    def initialize_parameters(input_dim: int, output_dim: int):
        dims["in"], dims["out"] = input_dim, output_dim
        torch.manual_seed(0)
        model = nn.Sequential(nn.Linear(input_dim, 16), nn.ReLU(), nn.Linear(16, output_dim))
        return ndarrays_to_parameters([v.detach() for v in model.state_dict().values()])

    strategy = BasicFedAvg(min_fit_clients=2, min_evaluate_clients=2, min_available_clients=2,
                           fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
                           evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn)
    schema = TabularFeaturesInfoEncoder.encoder_from_dataframe(_frame(0), "pid", "outcome") if server_has_schema else None
    config = {"n_server_rounds": 3, "batch_size": 16, "local_epochs": 1}
    server = TabularFeatureAlignmentServer(SimpleClientManager(), config, initialize_parameters, strategy, schema)
    history = run_simulation(server, clients, 3)
    assert dims["out"] == 2 and dims["in"] == clients[0].input_dimension == clients[1].input_dimension
    assert len(history.losses_distributed) == 3
    assert history.losses_distributed[-1][1] < history.losses_distributed[0][1]
