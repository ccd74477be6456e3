"""The JSON-serialisable alignment schema ("source of truth") exchanged between server and clients (parity:
``fl4health/feature_alignment/tab_features_info_encoder.py:14-125``)."""

from __future__ import annotations

import json

import pandas as pd

from fl4health_b200.common.typing import Scalar
from fl4health_b200.feature_alignment.feature_type_extraction import TabularFeatures
from fl4health_b200.feature_alignment.tabular_feature import MetaData, TabularFeature
from fl4health_b200.feature_alignment.tabular_type import TabularType


class TabularFeaturesInfoEncoder:
    def __init__(self, tabular_features: list[TabularFeature], tabular_targets: list[TabularFeature]) -> None:
        self.tabular_features = sorted(tabular_features, key=TabularFeature.get_feature_name)
        self.tabular_targets = sorted(tabular_targets, key=TabularFeature.get_feature_name)

    def get_tabular_features(self) -> list[TabularFeature]:
        return self.tabular_features

    def get_tabular_targets(self) -> list[TabularFeature]:
        return self.tabular_targets

    def get_feature_columns(self) -> list[str]:
        return sorted(f.get_feature_name() for f in self.tabular_features)

    def get_target_columns(self) -> list[str]:
        return sorted(t.get_feature_name() for t in self.tabular_targets)

    def features_by_type(self, tabular_type: TabularType) -> list[TabularFeature]:
        return [f for f in self.tabular_features if f.get_feature_type() == tabular_type]

    def type_to_features(self) -> dict[TabularType, list[TabularFeature]]:
        return {t: self.features_by_type(t) for t in TabularType}

    def get_categories_list(self) -> list[MetaData]:
        return [f.get_metadata() for f in self.features_by_type(TabularType.ORDINAL)]

    def get_target_dimension(self) -> int:
        return sum(t.get_metadata_dimension() for t in self.tabular_targets)

    @staticmethod
    def _construct_tab_feature(df: pd.DataFrame, feature_name: str, feature_type: TabularType,
                               fill_values: dict[str, Scalar] | None) -> TabularFeature:
        fill = (fill_values or {}).get(feature_name, TabularType.get_default_fill_value(feature_type))
        if feature_type in {TabularType.ORDINAL, TabularType.BINARY}:
            return TabularFeature(feature_name, feature_type, fill, sorted(df[feature_name].unique().tolist()))
        if feature_type == TabularType.STRING:
            from sklearn.feature_extraction.text import CountVectorizer

            vocabulary = CountVectorizer().fit(df[feature_name]).vocabulary_
            return TabularFeature(feature_name, feature_type, fill, {k: int(v) for k, v in vocabulary.items()})
        return TabularFeature(feature_name, feature_type, fill)

    @staticmethod
    def encoder_from_dataframe(df: pd.DataFrame, id_column: str, target_columns: str | list[str],
                               fill_values: dict[str, Scalar] | None = None) -> TabularFeaturesInfoEncoder:
        columns = sorted(c for c in df.columns.tolist() if c != id_column)
        types = TabularFeatures(data=df.reset_index(), features=columns, by=id_column, targets=target_columns).types
        targets_set = {target_columns} if isinstance(target_columns, str) else set(target_columns)
        features, targets = [], []
        for name, feature_type in types.items():
            feature = TabularFeaturesInfoEncoder._construct_tab_feature(df, name, TabularType(feature_type.value), fill_values)
            (targets if name in targets_set else features).append(feature)
        return TabularFeaturesInfoEncoder(features, targets)

    def to_json(self) -> str:
        return json.dumps({
            "tabular_features": json.dumps([f.to_json() for f in self.tabular_features]),
            "tabular_targets": json.dumps([t.to_json() for t in self.tabular_targets]),
        })

    @staticmethod
    def from_json(json_str: str) -> TabularFeaturesInfoEncoder:
        fields = json.loads(json_str)
        return TabularFeaturesInfoEncoder(
            [TabularFeature.from_json(s) for s in json.loads(fields["tabular_features"])],
            [TabularFeature.from_json(s) for s in json.loads(fields["tabular_targets"])],
        )
