"""One column of the alignment schema (parity: ``fl4health/feature_alignment/tabular_feature.py:13-98``)."""

from __future__ import annotations

import json

from fl4health_b200.common.typing import Scalar
from fl4health_b200.feature_alignment.tabular_type import TabularType

MetaData = dict[str, int] | list[Scalar]  # categories (binary / ordinal) or vocabulary (string)


class TabularFeature:
    def __init__(self, feature_name: str, feature_type: TabularType, fill_value: Scalar | None, metadata: MetaData | None = None) -> None:
        self.feature_name = feature_name
        self.feature_type = feature_type
        self.fill_value = TabularType.get_default_fill_value(feature_type) if fill_value is None else fill_value
        self.metadata: MetaData = metadata if metadata else []

    def get_feature_name(self) -> str:
        return self.feature_name

    def get_feature_type(self) -> TabularType:
        return self.feature_type

    def get_fill_value(self) -> Scalar:
        return self.fill_value

    def get_metadata(self) -> MetaData:
        return self.metadata

    def get_metadata_dimension(self) -> int:
        """Width this feature occupies as a TARGET after alignment."""
        if self.feature_type in {TabularType.BINARY, TabularType.ORDINAL}:
            return len(self.metadata)
        if self.feature_type == TabularType.NUMERIC:
            return 1
        raise ValueError("Metadata dimension is not supported when self.feature_type is TabularType.STRING.")

    def to_json(self) -> str:
        return json.dumps({
            "feature_name": json.dumps(self.feature_name), "feature_type": json.dumps(self.feature_type),
            "fill_value": json.dumps(self.fill_value), "metadata": json.dumps(self.metadata),
        })

    @staticmethod
    def from_json(json_str: str) -> TabularFeature:
        fields = json.loads(json_str)
        return TabularFeature(
            json.loads(fields["feature_name"]), TabularType(json.loads(fields["feature_type"])),
            json.loads(fields["fill_value"]), json.loads(fields["metadata"]),
        )
