"""sklearn transformers that apply a text vectoriser column-wise to DataFrames (parity:
``fl4health/feature_alignment/string_columns_transformer.py:9-88``)."""

from __future__ import annotations

from typing import Any

import pandas as pd
from sklearn.base import BaseEstimator, TransformerMixin


class TextMulticolumnTransformer(BaseEstimator, TransformerMixin):
    """Joins all string columns of a frame into one text per row, then vectorises."""

    def __init__(self, transformer: Any) -> None:
        self.transformer = transformer

    @staticmethod
    def _joined(x: pd.DataFrame) -> pd.Series:
        return x.astype(str).agg(" ".join, axis=1)

    def fit(self, x: pd.DataFrame, y: pd.DataFrame | None = None) -> TextMulticolumnTransformer:  # noqa: ARG002
        self.transformer.fit(self._joined(x))
        return self

    def transform(self, x: pd.DataFrame) -> pd.DataFrame:
        return self.transformer.transform(self._joined(x))


class TextColumnTransformer(BaseEstimator, TransformerMixin):
    """Vectorises a single-column frame (what ``ColumnTransformer`` hands over for one feature)."""

    def __init__(self, transformer: Any) -> None:
        self.transformer = transformer

    def fit(self, x: pd.DataFrame, y: pd.DataFrame | None = None) -> TextColumnTransformer:  # noqa: ARG002
        assert isinstance(x, pd.DataFrame) and x.shape[1] == 1
        self.transformer.fit(x.iloc[:, 0].astype(str))
        return self

    def transform(self, x: pd.DataFrame) -> pd.DataFrame:
        assert isinstance(x, pd.DataFrame) and x.shape[1] == 1
        return self.transformer.transform(x.iloc[:, 0].astype(str))
