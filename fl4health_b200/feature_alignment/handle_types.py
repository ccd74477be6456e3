"""Column type inference and conversion for tabular feature alignment.

Behavioural parity with ``fl4health/feature_alignment/handle_types.py`` (the 20 helpers at :22-587), organised as one
table: every ``FeatureType`` has a ``_Kind`` row holding *can this column become that type* and *make it so*.  The
public entry points (``infer_types``, ``to_types``, ``convertible_to_type``, ``to_dtype``, ``get_unique``,
``valid_feature_type``) and the reference's private helper names (``_to_ordinal``, ``_convertible_to_binary`` ...,
kept because downstream code of the reference imports them) are views on that table.

Inference order, first match wins (``handle_types.py:470-498``):

1. BINARY   bool dtype, or exactly two distinct non-null values (numeric columns only if integer typed);
2. ORDINAL  between 2 and ``ORDINAL_MAX_CATEGORIES`` (20) distinct values (again, not for float columns);
3. NUMERIC  accepted by ``pd.to_numeric``;
4. STRING   always possible.

Conversions return ``(data, metadata)``; categorical conversions record the inverse mapping (code -> original value)
under ``FEATURE_MAPPING_ATTR``, one-hot conversion records the source column under ``FEATURE_INDICATOR_ATTR``.
"""

from __future__ import annotations

from collections.abc import Callable
from dataclasses import dataclass
from typing import Any

import numpy as np
import pandas as pd
from pandas.api.types import is_bool_dtype, is_integer_dtype, is_numeric_dtype

from fl4health_b200.feature_alignment.constants import (
    FEATURE_INDICATOR_ATTR,
    FEATURE_MAPPING_ATTR,
    FEATURE_TYPE_ATTR,
    FEATURE_TYPES,
    ORDINAL_MAX_CATEGORIES,
    FeatureType,
)

Meta = dict[str, Any]


# ----------------------------------------------------------------------------------------------------------------------
# small shared pieces
# ----------------------------------------------------------------------------------------------------------------------
def valid_feature_type(type: FeatureType, raise_error: bool = True) -> bool:  # noqa: A002
    """Whether ``type`` is one of the feature types a column can be converted to (parity: handle_types.py:393-414)."""
    if type in FEATURE_TYPES:
        return True
    if raise_error:
        raise ValueError(f"Feature type '{type.value}' not in {', '.join(t.value for t in FEATURE_TYPES)}.")
    return False


def get_unique(values: np.ndarray | pd.Series, unique: np.ndarray | None = None) -> np.ndarray:
    """Distinct values of a column, unless the caller already has them (parity: handle_types.py:373-390)."""
    if unique is not None:
        return unique
    return values.unique() if isinstance(values, pd.Series) else pd.Series(values).unique()


def _distinct_non_null(series: pd.Series, unique: np.ndarray | None) -> int | None:
    """How many categories the column would have; None when it cannot be categorical at all (float columns)."""
    if is_numeric_dtype(series) and not is_integer_dtype(series):
        return None
    values = get_unique(series, unique)
    return int((~pd.isnull(values)).sum())


def _convertible_to_categorical(
    series: pd.Series,
    category_min: int | None = None,
    category_max: int | None = None,
    unique: np.ndarray | None = None,
    raise_error_over_max: bool = False,
    raise_error_under_min: bool = False,
) -> bool:
    """``category_min <= #categories <= category_max`` (open bounds when None); optionally an error instead of False
    when a bound is violated (parity: handle_types.py:271-326)."""
    count = _distinct_non_null(series, unique)
    if count is None:
        return False
    too_few = category_min is not None and count < category_min
    too_many = category_max is not None and count > category_max
    if too_many and raise_error_over_max:
        raise ValueError(f"Should have at most {category_max} categories, but has {count}.")
    if too_few and raise_error_under_min:
        raise ValueError(f"Should have at least {category_min} categories, but has {count}.")
    return not (too_few or too_many)


def _numeric_categorical_mapping(series: pd.Series, unique: np.ndarray | None = None) -> tuple[pd.Series, Meta]:
    """Replace the values by their rank among the sorted distinct values; the metadata holds rank -> value
    (parity: handle_types.py:152-181; object columns are ranked as strings)."""
    values = get_unique(series, unique)
    if values.dtype.name == "object":
        values = values.astype(str)
    ranks = {value: rank for rank, value in enumerate(np.sort(values))}
    return series.map(ranks), {FEATURE_MAPPING_ATTR: {rank: value for value, rank in ranks.items()}}


# ----------------------------------------------------------------------------------------------------------------------
# the table
# ----------------------------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class _Kind:
    dtype: str | None  # pandas dtype that goes with the type (None: keep what the column has)
    accepts: Callable[[pd.Series, np.ndarray | None], bool]
    convert: Callable[[pd.Series, np.ndarray | None], tuple[pd.Series, Meta]] | None  # None: frame-level conversion


def _numeric_ok(series: pd.Series, unique: np.ndarray | None = None) -> bool:
    try:
        pd.to_numeric(series)
    except (ValueError, TypeError):
        return False
    return True


def _as_numeric(series: pd.Series, unique: np.ndarray | None = None) -> tuple[pd.Series, Meta]:
    return pd.to_numeric(series), {FEATURE_TYPE_ATTR: FeatureType.NUMERIC}


def _as_string(series: pd.Series, unique: np.ndarray | None = None) -> tuple[pd.Series, Meta]:
    return series, {FEATURE_TYPE_ATTR: FeatureType.STRING}


def _as_ranked(feature_type: FeatureType) -> Callable[[pd.Series, np.ndarray | None], tuple[pd.Series, Meta]]:
    def convert(series: pd.Series, unique: np.ndarray | None = None) -> tuple[pd.Series, Meta]:
        if feature_type == FeatureType.BINARY and is_bool_dtype(series):
            return series.astype("category"), {FEATURE_TYPE_ATTR: feature_type, FEATURE_MAPPING_ATTR: {False: False, True: True}}
        ranked, meta = _numeric_categorical_mapping(series, unique)
        return ranked.astype("category"), {**meta, FEATURE_TYPE_ATTR: feature_type}

    return convert


def _bounded(low: int, high: int, bool_ok: bool = False) -> Callable[[pd.Series, np.ndarray | None], bool]:
    def accepts(series: pd.Series, unique: np.ndarray | None = None) -> bool:
        if bool_ok and is_bool_dtype(series):
            return True
        return _convertible_to_categorical(series, category_min=low, category_max=high, unique=unique)

    return accepts


_KINDS: dict[FeatureType, _Kind] = {
    FeatureType.BINARY: _Kind("category", _bounded(2, 2, bool_ok=True), _as_ranked(FeatureType.BINARY)),
    FeatureType.ORDINAL: _Kind("category", _bounded(2, ORDINAL_MAX_CATEGORIES), _as_ranked(FeatureType.ORDINAL)),
    FeatureType.CATEGORICAL_INDICATOR: _Kind("category", _bounded(2, ORDINAL_MAX_CATEGORIES), None),
    FeatureType.NUMERIC: _Kind(None, _numeric_ok, _as_numeric),
    FeatureType.STRING: _Kind(None, lambda series, unique=None: True, _as_string),
}
_INFERENCE_ORDER = (FeatureType.BINARY, FeatureType.ORDINAL, FeatureType.NUMERIC, FeatureType.STRING)


def _kind(type: FeatureType) -> _Kind:  # noqa: A002
    if type not in _KINDS:
        valid_feature_type(type, raise_error=True)  # unknown types fail here ...
        raise ValueError("Supported type has no corresponding datatype.")  # ... known ones without a rule, here
    return _KINDS[type]


# ----------------------------------------------------------------------------------------------------------------------
# public surface
# ----------------------------------------------------------------------------------------------------------------------
def _type_to_dtype(type: FeatureType) -> str | None:  # noqa: A002
    return _kind(type).dtype


def to_dtype(series: pd.Series, type: FeatureType) -> pd.Series:  # noqa: A002
    """``series`` with the pandas dtype that goes with the feature type (categorical for binary / ordinal / indicator
    columns, untouched otherwise; parity: handle_types.py:448-467)."""
    dtype = _type_to_dtype(type)
    return series if dtype is None or series.dtype == dtype else series.astype(dtype)


def convertible_to_type(series: pd.Series, type: FeatureType, unique: np.ndarray | None = None, raise_error: bool = False) -> bool:  # noqa: A002
    ok = bool(_kind(type).accepts(series, unique))
    if raise_error and not ok:
        raise ValueError(f"Cannot convert series {series.name} to type {type}.")
    return ok


def _infer_type(series: pd.Series, unique: np.ndarray | None = None) -> FeatureType:
    unique = get_unique(series, unique)
    for candidate in _INFERENCE_ORDER:
        if convertible_to_type(series, candidate, unique=unique):
            return candidate
    raise ValueError(f"Could not infer type of series '{series.name}'.")


def infer_types(data: pd.DataFrame, features: list[str]) -> dict[str, FeatureType]:
    return {name: _infer_type(data[name]) for name in features}


def _to_categorical_indicators(data: pd.DataFrame, col: str, unique: np.ndarray | None = None) -> tuple[pd.DataFrame, Meta]:
    """One-hot: ``col`` is replaced by one categorical indicator column per value, named ``{col}_{value}``; each
    records which column it came from (parity: handle_types.py:64-103)."""
    indicators = pd.get_dummies(data[col], prefix=str(data[col].name))
    clashes = set(indicators.columns) & set(data.columns)
    if clashes:
        raise ValueError(f"Cannot duplicate columns {', '.join(clashes)}.")
    indicators = indicators.apply(lambda column: to_dtype(column, FeatureType.CATEGORICAL_INDICATOR))
    meta = {name: {FEATURE_TYPE_ATTR: FeatureType.CATEGORICAL_INDICATOR, FEATURE_INDICATOR_ATTR: col} for name in indicators.columns}
    return pd.concat([data.drop(columns=[col]), indicators], axis=1), meta


def _to_type(data: pd.DataFrame, col: str, new_type: FeatureType, unique: np.ndarray | None = None) -> tuple[pd.DataFrame, Meta]:
    if data is None:
        raise ValueError("The features data must be passed to keyword argument 'data'.")
    kind = _kind(new_type)
    if kind.convert is None:
        return _to_categorical_indicators(data, col, unique=unique)
    convertible_to_type(data[col], new_type, unique=unique, raise_error=True)
    converted, meta = kind.convert(data[col], unique)
    data[col] = converted
    return data, {str(converted.name): meta}


def to_types(data: pd.DataFrame, new_types: dict[str, FeatureType]) -> tuple[pd.DataFrame, Meta]:
    """Convert the named columns (in place, like the reference); returns the frame and per-column metadata."""
    collected: Meta = {}
    for col, new_type in new_types.items():
        data, meta = _to_type(data, col, new_type)
        collected.update(meta)
    return data, collected


# ---- the reference's per-type helper names, as views on the table --------------------------------------------------
def _convertible_to_binary(series: pd.Series, unique: np.ndarray | None = None) -> bool:
    return convertible_to_type(series, FeatureType.BINARY, unique=unique)


def _convertible_to_numeric(series: pd.Series, raise_error: bool = False) -> bool:
    if raise_error:
        pd.to_numeric(series)  # pandas' own error says which value is the problem
        return True
    return _numeric_ok(series)


def _convertible_to_ordinal(series: pd.Series, unique: np.ndarray | None = None, category_max: int = ORDINAL_MAX_CATEGORIES,
                            raise_error_over_max: bool = False) -> bool:
    return _convertible_to_categorical(series, category_min=2, category_max=category_max, unique=unique,
                                       raise_error_over_max=raise_error_over_max)


_convertible_to_categorical_indicators = _convertible_to_ordinal  # same bounds: 2 .. category_max distinct values


def _to_binary(series: pd.Series, unique: np.ndarray | None = None) -> tuple[pd.Series, Meta]:
    return _KINDS[FeatureType.BINARY].convert(series, unique)  # type: ignore[misc]


def _to_ordinal(series: pd.Series, unique: np.ndarray | None = None) -> tuple[pd.Series, Meta]:
    return _KINDS[FeatureType.ORDINAL].convert(series, unique)  # type: ignore[misc]


def _to_numeric(series: pd.Series, unique: np.ndarray | None = None) -> tuple[pd.Series, Meta]:
    return _as_numeric(series, unique)


def _to_string(series: pd.Series) -> tuple[pd.Series, Meta]:
    return _as_string(series)
