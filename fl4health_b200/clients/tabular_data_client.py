"""Client for tabular data with server-coordinated feature alignment (parity:
``fl4health/clients/tabular_data_client.py:22-187``).

Protocol (driven by ``TabularFeatureAlignmentServer`` through ``get_properties`` polls before round 1):
  1. ``source_specified == False`` -> reply with this client's schema (JSON) — one client becomes the source of truth;
  2. ``source_specified == True``  -> align the local frame to the schema in ``config[feature_info]``, build loaders
     and model, and reply with the aligned input / output dimensions so the server can size the global model.
"""

from __future__ import annotations

from collections.abc import Sequence
from logging import INFO
from pathlib import Path
from typing import Any

import numpy as np
import torch

from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config, Scalar
from fl4health_b200.feature_alignment.constants import FEATURE_INFO, INPUT_DIMENSION, OUTPUT_DIMENSION, SOURCE_SPECIFIED
from fl4health_b200.feature_alignment.tab_features_info_encoder import TabularFeaturesInfoEncoder
from fl4health_b200.feature_alignment.tab_features_preprocessor import TabularFeaturesPreprocessor
from fl4health_b200.metrics.base_metrics import Metric
from fl4health_b200.utils.config import narrow_dict_type


class TabularDataClient(BasicClient):
    def __init__(self, data_path: Path, metrics: Sequence[Metric], device: torch.device, id_column: str,
                 targets: str | list[str], **basic_client_options: Any) -> None:
        """``id_column`` / ``targets`` name the bookkeeping columns of the local frame; everything else is
        ``BasicClient``'s (``loss_meter_type``, ``checkpoint_and_state_module``, ``reporters``, ...)."""
        super().__init__(data_path, metrics, device, **basic_client_options)
        self.id_column, self.targets = id_column, targets
        self.feature_specific_pipelines: dict[str, Any] = {}
        self.tabular_features_info_encoder: TabularFeaturesInfoEncoder
        self.tabular_features_preprocessor: TabularFeaturesPreprocessor
        self.df: Any
        self.aligned_features: np.ndarray
        self.aligned_targets: np.ndarray
        self.input_dimension: int
        self.output_dimension: int

    # ---------------------------------------------------------------------------------------------- user hooks
    def get_data_frame(self, config: Config) -> Any:
        """User hook: the local ``pandas.DataFrame`` (must contain ``id_column`` and the target column(s))."""
        raise NotImplementedError

    def preset_specific_pipeline(self, feature_name: str, pipeline: Any) -> None:
        """Override the default sklearn pipeline of one feature (call before the alignment poll)."""
        self.feature_specific_pipelines[feature_name] = pipeline

    def set_feature_specific_pipelines(self) -> None:
        for feature_name, pipeline in self.feature_specific_pipelines.items():
            self.tabular_features_preprocessor.set_feature_pipeline(feature_name, pipeline)

    # ---------------------------------------------------------------------------------------------- the two polls
    def _describe_local_schema(self) -> None:
        """Poll 1: summarise the local frame (it may be elected the federation's source of truth)."""
        self.tabular_features_info_encoder = TabularFeaturesInfoEncoder.encoder_from_dataframe(self.df, self.id_column, self.targets)

    def _align_to(self, schema_json: str) -> None:
        """Poll 2: adopt the agreed schema and turn the frame into aligned feature / target matrices."""
        self.tabular_features_info_encoder = TabularFeaturesInfoEncoder.from_json(schema_json)
        self.tabular_features_preprocessor = TabularFeaturesPreprocessor(self.tabular_features_info_encoder)
        self.set_feature_specific_pipelines()
        features, self.aligned_targets = self.tabular_features_preprocessor.preprocess_features(self.df)
        self.aligned_features = features.toarray() if hasattr(features, "toarray") else features  # one-hot / TF-IDF blocks are sparse
        self.input_dimension = self.aligned_features.shape[1]
        self.output_dimension = self.tabular_features_info_encoder.get_target_dimension()
        log(INFO, f"input dimension: {self.input_dimension}, target dimension: {self.output_dimension}")

    def setup_client(self, config: Config) -> None:
        self.df = self.get_data_frame(config)
        if not narrow_dict_type(config, SOURCE_SPECIFIED, bool):
            self._describe_local_schema()
            return  # not "initialized": the real set-up happens once a schema has been agreed
        self._align_to(narrow_dict_type(config, FEATURE_INFO, str))
        super().setup_client(config)
        del self.aligned_features, self.aligned_targets, self.df  # the loaders own the tensors now

    def get_properties(self, config: Config) -> dict[str, Scalar]:
        if not self.initialized:
            self.setup_client(config)
        if narrow_dict_type(config, SOURCE_SPECIFIED, bool):
            return {INPUT_DIMENSION: self.input_dimension, OUTPUT_DIMENSION: self.output_dimension}
        return {FEATURE_INFO: self.tabular_features_info_encoder.to_json()}
