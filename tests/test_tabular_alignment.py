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


def test_type_conversion_helpers_cover_every_feature_type() -> None:
    """The conversion table behind ``infer_types`` / ``to_types`` (reference handle_types.py: 20 helpers): ranking for
    binary / ordinal columns with the inverse map in the metadata, one-hot for categorical indicators, bounds on the
    number of categories, and floats never being categorical."""
    import numpy as np
    import pandas as pd
    import pytest

    from fl4health_b200.feature_alignment import handle_types as ht
    from fl4health_b200.feature_alignment.constants import FEATURE_INDICATOR_ATTR, FEATURE_MAPPING_ATTR, FEATURE_TYPE_ATTR, FeatureType

    frame = pd.DataFrame({"flag": [True, False, True, True], "yn": ["y", "n", "y", "n"], "grade": [3, 1, 2, 3],
                          "dose": [0.5, 1.5, 2.5, 0.5], "colour": ["r", "g", "b", "r"]})
    assert ht.infer_types(frame, list(frame.columns)) == {
        "flag": FeatureType.BINARY, "yn": FeatureType.BINARY, "grade": FeatureType.ORDINAL, "dose": FeatureType.NUMERIC,
        "colour": FeatureType.ORDINAL}
    wide = pd.DataFrame({"name": [f"p{i}" for i in range(30)], "digits": [str(i) for i in range(30)]})
    assert ht.infer_types(wide, ["name", "digits"]) == {"name": FeatureType.STRING, "digits": FeatureType.NUMERIC}

    converted, meta = ht.to_types(frame.copy(), {"flag": FeatureType.BINARY, "yn": FeatureType.BINARY, "grade": FeatureType.ORDINAL,
                                                 "dose": FeatureType.NUMERIC, "colour": FeatureType.CATEGORICAL_INDICATOR})
    assert converted["yn"].tolist() == [1, 0, 1, 0] and meta["yn"][FEATURE_MAPPING_ATTR] == {0: "n", 1: "y"}
    assert converted["grade"].tolist() == [2, 0, 1, 2] and meta["grade"][FEATURE_MAPPING_ATTR] == {0: 1, 1: 2, 2: 3}
    assert meta["flag"][FEATURE_MAPPING_ATTR] == {False: False, True: True} and str(converted["flag"].dtype) == "category"
    assert "colour" not in converted and {"colour_r", "colour_g", "colour_b"} <= set(converted.columns)
    assert converted["colour_r"].tolist() == [True, False, False, True]
    assert meta["colour_g"] == {FEATURE_TYPE_ATTR: FeatureType.CATEGORICAL_INDICATOR, FEATURE_INDICATOR_ATTR: "colour"}
    assert meta["dose"] == {FEATURE_TYPE_ATTR: FeatureType.NUMERIC}

    # bounds and errors
    assert not ht.convertible_to_type(frame["dose"], FeatureType.ORDINAL)  # floats are never categorical
    assert not ht._convertible_to_ordinal(wide["name"]) and ht._convertible_to_ordinal(wide["name"], category_max=30)
    with pytest.raises(ValueError, match="at most 20"):
        ht._convertible_to_ordinal(wide["name"], raise_error_over_max=True)
    with pytest.raises(ValueError, match="at least 2"):
        ht._convertible_to_categorical(pd.Series([1, 1, 1]), category_min=2, raise_error_under_min=True)
    with pytest.raises(ValueError, match="Cannot convert series dose"):
        ht.to_types(frame.copy(), {"dose": FeatureType.BINARY})
    with pytest.raises(ValueError, match="Cannot duplicate columns"):
        ht._to_categorical_indicators(pd.DataFrame({"c": ["a", "b"], "c_a": [0, 1]}), "c")
    assert ht._convertible_to_numeric(pd.Series(["1", "2"])) and not ht._convertible_to_numeric(pd.Series(["1", "x"]))
    with pytest.raises(ValueError):
        ht._convertible_to_numeric(pd.Series(["1", "x"]), raise_error=True)
    # caller-supplied distinct values are honoured, missing values are not a category
    assert ht._convertible_to_binary(pd.Series(["a", None, "b", "a"]))
    ranked, ranked_meta = ht._to_ordinal(pd.Series(["lo", "hi", "mid"], name="level"), unique=np.array(["hi", "lo", "mid"], dtype=object))
    assert ranked.tolist() == [1, 0, 2] and ranked_meta[FEATURE_TYPE_ATTR] == FeatureType.ORDINAL
    assert ht.to_dtype(pd.Series([0, 1]), FeatureType.NUMERIC).dtype == np.int64 and ht._type_to_dtype(FeatureType.STRING) is None
