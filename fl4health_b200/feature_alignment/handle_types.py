"""Column type inference and conversion (condensed equivalent of ``fl4health/feature_alignment/handle_types.py``).

Inference order (first match wins), as in the reference (``handle_types.py:470-498``):

1. BINARY   bool dtype, or exactly two distinct non-null values (numeric columns only if integer typed);
2. ORDINAL  between 2 and ``ORDINAL_MAX_CATEGORIES`` (20) distinct values (again, not for float columns);
3. NUMERIC  convertible by ``pd.to_numeric``;
4. STRING   anything else made of strings.
"""

from __future__ import annotations

from typing import Any

import numpy as np
import pandas as pd
from pandas.api.types import is_bool_dtype, is_integer_dtype, is_numeric_dtype, is_object_dtype, is_string_dtype

from fl4health_b200.feature_alignment.constants import (
    FEATURE_TYPES,
    FEATURE_MAPPING_ATTR,
    FEATURE_TYPE_ATTR,
    ORDINAL_MAX_CATEGORIES,
    FeatureType,
)


def valid_feature_type(type: FeatureType, raise_error: bool = True) -> bool:  # noqa: A002
    """Whether ``type`` is one of the feature types a column can be converted to (parity: handle_types.py:393-414)."""
    if type in FEATURE_TYPES:
        return True
    if raise_error:
        raise ValueError(f"Feature type '{type.value}' not in {', '.join(t.value for t in FEATURE_TYPES)}.")
    return False


def _type_to_dtype(type: FeatureType) -> str | None:  # noqa: A002
    if type in (FeatureType.STRING, FeatureType.NUMERIC):
        return None  # the caller keeps its own string length / numeric precision
    if type in (FeatureType.BINARY, FeatureType.CATEGORICAL_INDICATOR, FeatureType.ORDINAL):
        return "category"
    if valid_feature_type(type, raise_error=True):
        raise ValueError("Supported type has no corresponding datatype.")
    return None


def to_dtype(series: pd.Series, type: FeatureType) -> pd.Series:  # noqa: A002
    """``series`` with the pandas dtype that goes with the feature type (categorical for binary / ordinal / indicator
    columns, untouched otherwise; parity: handle_types.py:448-467)."""
    dtype = _type_to_dtype(type)
    if dtype is None or series.dtype == dtype:
        return series
    return series.astype(dtype)


def get_unique(values: np.ndarray | pd.Series, unique: np.ndarray | None = None) -> np.ndarray:
    if unique is not None:
        return unique
    return pd.Series(values).unique() if not isinstance(values, pd.Series) else values.unique()


def _n_categories(series: pd.Series, unique: np.ndarray | None) -> int | None:
    """Distinct non-null values, or None when the column cannot be categorical (non-integer numerics)."""
    if is_numeric_dtype(series) and not is_integer_dtype(series) and not is_bool_dtype(series):
        return None
    uniq = get_unique(series, unique)
    return int((~pd.isnull(uniq)).sum())


def convertible_to_type(series: pd.Series, type: FeatureType, unique: np.ndarray | None = None, raise_error: bool = False) -> bool:  # noqa: A002
    ok: bool
    if type == FeatureType.BINARY:
        ok = bool(is_bool_dtype(series)) or _n_categories(series, unique) == 2
    elif type == FeatureType.ORDINAL:
        count = _n_categories(series, unique)
        ok = count is not None and 2 <= count <= ORDINAL_MAX_CATEGORIES
    elif type == FeatureType.NUMERIC:
        try:
            pd.to_numeric(series)
            ok = True
        except (ValueError, TypeError):
            ok = False
    elif type == FeatureType.STRING:
        ok = bool(is_string_dtype(series) or is_object_dtype(series))
    else:
        raise ValueError(f"Unsupported feature type {type}")
    if not ok and raise_error:
        raise ValueError(f"Cannot convert series '{series.name}' to type {type.value}.")
    return ok


def _infer_type(series: pd.Series, unique: np.ndarray | None = None) -> FeatureType:
    unique = get_unique(series, unique)
    for candidate in (FeatureType.BINARY, FeatureType.ORDINAL, FeatureType.NUMERIC, FeatureType.STRING):
        if convertible_to_type(series, candidate, unique=unique):
            return candidate
    raise ValueError(f"Could not infer type of series '{series.name}'.")


def infer_types(data: pd.DataFrame, features: list[str]) -> dict[str, FeatureType]:
    return {name: _infer_type(data[name]) for name in features}


def _category_mapping(series: pd.Series) -> tuple[pd.Series, dict[Any, int]]:
    categories = sorted(series.dropna().unique().tolist(), key=lambda v: (str(type(v)), v))
    mapping = {value: index for index, value in enumerate(categories)}
    return series.map(mapping), mapping


def to_types(data: pd.DataFrame, new_types: dict[str, FeatureType]) -> tuple[pd.DataFrame, dict[str, Any]]:
    """Convert columns to the requested types; returns the new frame and per-column metadata (type + category mapping)."""
    out = data.copy()
    meta: dict[str, Any] = {}
    for name, feature_type in new_types.items():
        convertible_to_type(out[name], feature_type, raise_error=True)
        if feature_type == FeatureType.NUMERIC:
            out[name] = pd.to_numeric(out[name]).astype(float)
            meta[name] = {FEATURE_TYPE_ATTR: feature_type}
        elif feature_type in (FeatureType.BINARY, FeatureType.ORDINAL):
            if is_bool_dtype(out[name]):
                mapping: dict[Any, Any] = {False: False, True: True}
            else:
                out[name], mapping = _category_mapping(out[name])
            meta[name] = {FEATURE_TYPE_ATTR: feature_type, FEATURE_MAPPING_ATTR: {v: k for k, v in mapping.items()}}
        else:
            out[name] = out[name].astype("string")
            meta[name] = {FEATURE_TYPE_ATTR: feature_type}
    return out, meta
