"""Schema-driven sklearn pipelines that give every client identically shaped arrays (parity:
``fl4health/feature_alignment/tab_features_preprocessor.py:18-222``).

Defaults: numeric -> mean-impute + min-max scale; binary -> most-frequent impute + ordinal encode; ordinal -> one-hot
(features) / ordinal code (targets) over the SCHEMA's category list; string -> TF-IDF over the SCHEMA's vocabulary.
Columns a client lacks are filled with the feature's fill value first."""

from __future__ import annotations

from logging import WARNING

import numpy as np
import pandas as pd
from sklearn.compose import ColumnTransformer
from sklearn.feature_extraction.text import TfidfVectorizer
from sklearn.impute import SimpleImputer
from sklearn.pipeline import Pipeline
from sklearn.preprocessing import MinMaxScaler, OneHotEncoder, OrdinalEncoder

from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Scalar
from fl4health_b200.feature_alignment.string_columns_transformer import TextColumnTransformer
from fl4health_b200.feature_alignment.tab_features_info_encoder import TabularFeaturesInfoEncoder
from fl4health_b200.feature_alignment.tabular_feature import MetaData, TabularFeature
from fl4health_b200.feature_alignment.tabular_type import TabularType


class TabularFeaturesPreprocessor:
    def __init__(self, tab_feature_encoder: TabularFeaturesInfoEncoder) -> None:
        self.tabular_features = tab_feature_encoder.get_tabular_features()
        self.tabular_targets = tab_feature_encoder.get_tabular_targets()
        self.feature_columns = tab_feature_encoder.get_feature_columns()
        self.target_columns = tab_feature_encoder.get_target_columns()
        self.features_to_pipelines = self.initialize_default_pipelines(self.tabular_features, one_hot=True)
        self.targets_to_pipelines = self.initialize_default_pipelines(self.tabular_targets, one_hot=False)
        self.data_column_transformer = self.return_column_transformer(self.features_to_pipelines)
        self.target_column_transformer = self.return_column_transformer(self.targets_to_pipelines)

    def get_default_numeric_pipeline(self) -> Pipeline:
        return Pipeline([("imputer", SimpleImputer(strategy="mean")), ("scaler", MinMaxScaler())])

    def get_default_binary_pipeline(self) -> Pipeline:
        return Pipeline([("imputer", SimpleImputer(strategy="most_frequent")), ("encoder", OrdinalEncoder())])

    def get_default_one_hot_pipeline(self, categories: MetaData) -> Pipeline:
        return Pipeline([("encoder", OneHotEncoder(handle_unknown="ignore", categories=[categories]))])

    def get_default_ordinal_pipeline(self, categories: MetaData) -> Pipeline:
        encoder = OrdinalEncoder(unknown_value=len(categories) + 1, handle_unknown="use_encoded_value", categories=[categories])
        return Pipeline([("encoder", encoder)])

    def get_default_string_pipeline(self, vocabulary: MetaData) -> Pipeline:
        return Pipeline([("vectorizer", TextColumnTransformer(TfidfVectorizer(vocabulary=vocabulary)))])

    def initialize_default_pipelines(self, tabular_features: list[TabularFeature], one_hot: bool) -> dict[str, Pipeline]:
        pipelines = {}
        for feature in tabular_features:
            kind = feature.get_feature_type()
            if kind == TabularType.NUMERIC:
                pipeline = self.get_default_numeric_pipeline()
            elif kind == TabularType.BINARY:
                pipeline = self.get_default_binary_pipeline()
            elif kind == TabularType.ORDINAL:
                make = self.get_default_one_hot_pipeline if one_hot else self.get_default_ordinal_pipeline
                pipeline = make(feature.get_metadata())
            else:
                pipeline = self.get_default_string_pipeline(feature.get_metadata())
            pipelines[feature.get_feature_name()] = pipeline
        return pipelines

    def return_column_transformer(self, pipelines: dict[str, Pipeline]) -> ColumnTransformer:
        transformers = [(f"{name}_pipeline", pipelines[name], [name]) for name in sorted(pipelines)]
        return ColumnTransformer(transformers=transformers, remainder="drop")  # unlisted columns are dropped

    def set_feature_pipeline(self, feature_name: str, pipeline: Pipeline) -> None:
        if feature_name in self.features_to_pipelines:
            self.features_to_pipelines[feature_name] = pipeline
            self.data_column_transformer = self.return_column_transformer(self.features_to_pipelines)
        elif feature_name in self.targets_to_pipelines:
            self.targets_to_pipelines[feature_name] = pipeline
            self.target_column_transformer = self.return_column_transformer(self.targets_to_pipelines)
        else:
            log(WARNING, f"{feature_name} is neither a feature nor target and the provided pipeline will be ignored.")

    def preprocess_features(self, df: pd.DataFrame) -> tuple[np.ndarray, np.ndarray]:
        filled = self.fill_in_missing_columns(df)
        return (
            self.data_column_transformer.fit_transform(filled[self.feature_columns]),
            self.target_column_transformer.fit_transform(filled[self.target_columns]),
        )

    def fill_in_missing_columns(self, df: pd.DataFrame) -> pd.DataFrame:
        out = df.copy(deep=True)
        for feature in self.tabular_features:
            self._fill_in_missing_column(out, feature.get_feature_name(), feature.get_fill_value())
        for column in out.columns:  # sklearn imputers / encoders reject the bool dtype
            if out[column].dtype == bool:
                out[column] = out[column].astype(int)
        return out

    def _fill_in_missing_column(self, df: pd.DataFrame, column_name: str, value: Scalar) -> None:
        if column_name not in df.columns:
            df[column_name] = value
