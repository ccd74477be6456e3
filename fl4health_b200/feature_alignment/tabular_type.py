"""Parity: ``fl4health/feature_alignment/tabular_type.py``."""

from __future__ import annotations

from enum import Enum

from fl4health_b200.common.typing import Scalar


class TabularType(str, Enum):
    NUMERIC = "numeric"
    BINARY = "binary"
    ORDINAL = "ordinal"
    STRING = "string"

    @staticmethod
    def get_default_fill_value(tabular_type: TabularType | str) -> Scalar:
        """Imputation value used when a client lacks the column altogether."""
        defaults: dict[str, Scalar] = {"numeric": 0.0, "binary": 0, "string": "N/A", "ordinal": "UNKNOWN"}
        key = tabular_type.value if isinstance(tabular_type, TabularType) else tabular_type
        if key not in defaults:
            raise ValueError("Invalid Tabular Data Type.")
        return defaults[key]
