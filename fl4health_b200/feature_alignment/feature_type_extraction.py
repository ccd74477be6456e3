"""Feature-type bookkeeping over a DataFrame (condensed equivalent of
``fl4health/feature_alignment/feature_type_extraction.py:18-284``)."""

from __future__ import annotations

from typing import Any

import pandas as pd

from fl4health_b200.feature_alignment.constants import FEATURE_TARGET_ATTR, FEATURE_TYPE_ATTR, FeatureType
from fl4health_b200.feature_alignment.handle_types import infer_types


def to_list(obj: Any) -> list[Any]:
    if isinstance(obj, list):
        return obj
    if isinstance(obj, (tuple, set)):
        return list(obj)
    return [obj]


def has_columns(data: pd.DataFrame, cols: str | list[str], exactly: bool = False, raise_error: bool = False) -> bool:
    wanted, present = set(to_list(cols)), set(data.columns)
    ok = wanted == present if exactly else wanted.issubset(present)
    if not ok and raise_error:
        raise ValueError(f"Columns {sorted(wanted - present)} are missing from the data frame.")
    return ok


class FeatureMeta:
    def __init__(self, **kwargs: Any) -> None:
        if FEATURE_TYPE_ATTR not in kwargs:
            raise ValueError("Must specify feature type.")
        self.type_: FeatureType = kwargs[FEATURE_TYPE_ATTR]
        self.target: bool = kwargs.get(FEATURE_TARGET_ATTR, False)

    def get_type(self) -> FeatureType:
        return self.type_

    def update(self, meta: list[tuple[str, Any]]) -> None:
        """Set several meta attributes at once: ``[(attribute name, value), ...]`` (feature_type_extraction.py:107-115)."""
        for name, value in meta:
            setattr(self, name, value)


class Features:
    def __init__(self, data: pd.DataFrame, features: str | list[str], by: str | list[str] | None,
                 targets: str | list[str] | None = None, force_types: dict[str, FeatureType] | None = None) -> None:
        self.data = data
        self.by = to_list(by) if by is not None else []
        self.features = [f for f in to_list(features) if f not in self.by]
        self.targets = to_list(targets) if targets is not None else []
        has_columns(data, self.features + self.by, raise_error=True)
        inferred = infer_types(data, self.features)
        inferred.update(force_types or {})
        self.meta = {name: FeatureMeta(**{FEATURE_TYPE_ATTR: t, FEATURE_TARGET_ATTR: name in self.targets}) for name, t in inferred.items()}

    @property
    def types(self) -> dict[str, FeatureType]:
        return {name: meta.get_type() for name, meta in self.meta.items()}


class TabularFeatures(Features):
    """One row per entity identified by ``by`` (a single id column)."""

    def __init__(self, data: pd.DataFrame, features: str | list[str], by: str, targets: str | list[str] | None = None,
                 force_types: dict[str, FeatureType] | None = None) -> None:
        if not isinstance(by, str):
            raise ValueError("Tabular features index input as a string representing a column.")
        super().__init__(data, features, by, targets, force_types)
