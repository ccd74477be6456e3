"""sklearn transformers that hand text columns of a DataFrame to a text vectoriser (parity:
``fl4health/feature_alignment/string_columns_transformer.py:9-88``).

Both public classes are the same adapter -- turn the frame into ONE series of strings, then ``fit`` / ``transform`` the
wrapped vectoriser on it -- and differ only in how the series is formed: all columns joined per row, or the single
column a ``ColumnTransformer`` passes for one feature."""

from __future__ import annotations

from typing import Any

import pandas as pd
from sklearn.base import BaseEstimator, TransformerMixin


class _TextFrameAdapter(BaseEstimator, TransformerMixin):
    def __init__(self, transformer: Any) -> None:
        self.transformer = transformer

    def _as_text(self, frame: pd.DataFrame) -> pd.Series:
        raise NotImplementedError

    def fit(self, x: pd.DataFrame, y: pd.DataFrame | None = None) -> Any:  # noqa: ARG002
        self.transformer.fit(self._as_text(x))
        return self

    def transform(self, x: pd.DataFrame) -> Any:
        """The vectoriser's output (a sparse matrix for the usual count / tf-idf vectorisers)."""
        return self.transformer.transform(self._as_text(x))


class TextMulticolumnTransformer(_TextFrameAdapter):
    """Every row's string columns joined with spaces into one document."""

    def _as_text(self, frame: pd.DataFrame) -> pd.Series:
        return frame.astype(str).agg(" ".join, axis=1)


class TextColumnTransformer(_TextFrameAdapter):
    """A single-column frame: that column's strings are the documents."""

    def _as_text(self, frame: pd.DataFrame) -> pd.Series:
        assert isinstance(frame, pd.DataFrame) and frame.shape[1] == 1, "expects exactly one text column"
        return frame.iloc[:, 0].astype(str)
