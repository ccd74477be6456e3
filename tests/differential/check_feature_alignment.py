"""Tabular feature alignment vs the reference on the same data frames: inferred column types, the JSON schema a client
sends to the server, the aligned arrays after preprocessing (also with columns missing on one site), the text column
transformers, and every type-conversion helper."""
import json

import numpy as np
import pandas as pd
from sklearn.feature_extraction.text import CountVectorizer, TfidfVectorizer

import fl4health.feature_alignment.handle_types as ref_types
import fl4health.feature_alignment.string_columns_transformer as ref_text
import fl4health.feature_alignment.tab_features_info_encoder as ref_enc
import fl4health.feature_alignment.tab_features_preprocessor as ref_pre
import fl4health_b200.feature_alignment.handle_types as my_types
import fl4health_b200.feature_alignment.string_columns_transformer as my_text
import fl4health_b200.feature_alignment.tab_features_info_encoder as my_enc
import fl4health_b200.feature_alignment.tab_features_preprocessor as my_pre

rng = np.random.default_rng(31)
agreed = 0


def hospital_frame(n: int, drop: tuple[str, ...] = (), extra_category: bool = False) -> pd.DataFrame:
    frame = pd.DataFrame({
        "patient": np.arange(n),
        "age": rng.integers(18, 90, n).astype(float),
        "lactate": rng.normal(2.0, 0.7, n),
        "sex": rng.choice(["F", "M"], n),
        "smoker": rng.choice([0, 1], n),
        "ward": rng.choice(["icu", "er", "surgery", "medicine"] + (["oncology"] if extra_category else []), n),
        "triage": rng.integers(1, 6, n),
        "note": rng.choice(["stable overnight", "fever and cough", "chest pain on arrival", "post operative day two"], n),
        "outcome": rng.choice(["home", "ward", "icu"], n),
    })
    frame.loc[rng.choice(n, n // 10, replace=False), "lactate"] = np.nan
    return frame.drop(columns=list(drop))


def dense(x) -> np.ndarray:
    return np.asarray(x.todense()) if hasattr(x, "todense") else np.asarray(x)


# -- schema extraction and JSON ---------------------------------------------------------------------------------------
source = hospital_frame(300)
for targets in ("outcome", ["outcome", "smoker"]):
    theirs = ref_enc.TabularFeaturesInfoEncoder.encoder_from_dataframe(source.copy(), "patient", targets)
    ours = my_enc.TabularFeaturesInfoEncoder.encoder_from_dataframe(source.copy(), "patient", targets)
    assert theirs.get_feature_columns() == ours.get_feature_columns() and theirs.get_target_columns() == ours.get_target_columns()
    assert theirs.get_target_dimension() == ours.get_target_dimension()
    for a, b in zip(theirs.get_tabular_features() + theirs.get_tabular_targets(), ours.get_tabular_features() + ours.get_tabular_targets()):
        assert a.get_feature_name() == b.get_feature_name() and a.get_feature_type().value == b.get_feature_type().value
        assert list(a.get_metadata()) == list(b.get_metadata()) and a.get_metadata_dimension() == b.get_metadata_dimension()
        assert a.get_fill_value() == b.get_fill_value() or (pd.isna(a.get_fill_value()) and pd.isna(b.get_fill_value()))
    assert json.loads(theirs.to_json()) == json.loads(ours.to_json())
    # each side reads the other's wire format
    assert json.loads(ref_enc.TabularFeaturesInfoEncoder.from_json(ours.to_json()).to_json()) == json.loads(my_enc.TabularFeaturesInfoEncoder.from_json(theirs.to_json()).to_json())
    agreed += 1

# -- alignment: another site with an unseen category and missing columns is mapped into the same feature space ------------
schema_ref = ref_enc.TabularFeaturesInfoEncoder.encoder_from_dataframe(source.copy(), "patient", "outcome")
schema_mine = my_enc.TabularFeaturesInfoEncoder.from_json(schema_ref.to_json())
for site in (hospital_frame(120), hospital_frame(80, drop=("triage",)), hospital_frame(90, drop=("note", "sex"), extra_category=True)):
    x_ref, y_ref = ref_pre.TabularFeaturesPreprocessor(schema_ref).preprocess_features(site.copy())
    x_mine, y_mine = my_pre.TabularFeaturesPreprocessor(schema_mine).preprocess_features(site.copy())
    x_ref, x_mine, y_ref, y_mine = dense(x_ref), dense(x_mine), dense(y_ref), dense(y_mine)
    assert x_ref.shape == x_mine.shape and y_ref.shape == y_mine.shape, (x_ref.shape, x_mine.shape)
    assert np.allclose(x_ref.astype(float), x_mine.astype(float), atol=1e-9, equal_nan=True)
    assert np.allclose(y_ref.astype(float), y_mine.astype(float), atol=1e-9)
    agreed += 1

# -- text transformers ---------------------------------------------------------------------------------------------------
notes = source[["note"]].copy()
notes["second"] = notes["note"].str.upper()
for make in (lambda: CountVectorizer(), lambda: TfidfVectorizer()):
    a = ref_text.TextMulticolumnTransformer(make()).fit(notes).transform(notes)
    b = my_text.TextMulticolumnTransformer(make()).fit(notes).transform(notes)
    assert type(a).__name__ == type(b).__name__ and np.allclose(dense(a).astype(float), dense(b).astype(float))
    a = ref_text.TextColumnTransformer(make()).fit(notes[["note"]]).transform(notes[["note"]])
    b = my_text.TextColumnTransformer(make()).fit(notes[["note"]]).transform(notes[["note"]])
    assert type(a).__name__ == type(b).__name__ and np.allclose(dense(a).astype(float), dense(b).astype(float))  # sparse in, sparse out
    agreed += 1

# -- type helpers --------------------------------------------------------------------------------------------------------
columns = {
    "numeric": pd.Series(rng.normal(size=40)), "numeric_text": pd.Series([str(v) for v in rng.integers(0, 100, 40)]),
    "binary_text": pd.Series(rng.choice(["yes", "no"], 40)), "binary_number": pd.Series(rng.choice([0.0, 1.0], 40)),
    "few_ints": pd.Series(rng.integers(0, 4, 40)), "few_words": pd.Series(rng.choice(["a", "b", "c"], 40)),
    "many_words": pd.Series([f"w{i}" for i in range(40)]), "with_nan": pd.Series([1.0, np.nan, 0.0, 1.0] * 10),
    "bools": pd.Series(rng.choice([True, False], 40)),
}
inferred_ref = ref_types.infer_types(pd.DataFrame(columns), list(columns))
inferred_mine = my_types.infer_types(pd.DataFrame(columns), list(columns))
assert {k: v.value for k, v in inferred_ref.items()} == {k: v.value for k, v in inferred_mine.items()}, (inferred_ref, inferred_mine)
agreed += 1
for name, series in columns.items():
    assert ref_types._convertible_to_numeric(series) == my_types._convertible_to_numeric(series), name
    assert ref_types._convertible_to_binary(series) == my_types._convertible_to_binary(series), name
    assert ref_types._convertible_to_ordinal(series) == my_types._convertible_to_ordinal(series), name
    assert ref_types._convertible_to_categorical(series) == my_types._convertible_to_categorical(series), name
    assert list(ref_types.get_unique(series)) == list(my_types.get_unique(series)) or all(
        (x == y) or (pd.isna(x) and pd.isna(y)) for x, y in zip(ref_types.get_unique(series), my_types.get_unique(series))), name
    for kind in ("NUMERIC", "BINARY", "STRING", "ORDINAL", "CATEGORICAL_INDICATOR"):
        type_ref, type_mine = getattr(ref_types.FeatureType, kind), getattr(my_types.FeatureType, kind)
        possible_ref, possible_mine = ref_types.convertible_to_type(series, type_ref), my_types.convertible_to_type(series, type_mine)
        assert possible_ref == possible_mine, (name, kind)
        if possible_ref:
            (out_ref, meta_ref), (out_mine, meta_mine) = (ref_types._to_type(pd.DataFrame({name: series.copy()}), name, type_ref),
                                                          my_types._to_type(pd.DataFrame({name: series.copy()}), name, type_mine))
            assert list(out_ref.columns) == list(out_mine.columns), (name, kind, list(out_ref.columns), list(out_mine.columns))
            for column in out_ref.columns:
                left, right = out_ref[column].to_numpy(dtype=object), out_mine[column].to_numpy(dtype=object)
                assert all((x == y) or (pd.isna(x) and pd.isna(y)) for x, y in zip(left, right)), (name, kind, column)
            normal = lambda meta: json.dumps(meta, default=lambda v: getattr(v, "value", str(v)), sort_keys=True)  # noqa: E731
            assert normal(meta_ref) == normal(meta_mine), (name, kind, meta_ref, meta_mine)
    agreed += 1
frame = pd.DataFrame({k: columns[k] for k in ("few_words", "binary_text", "numeric_text")})
wanted_ref = {"few_words": ref_types.FeatureType.CATEGORICAL_INDICATOR, "binary_text": ref_types.FeatureType.BINARY, "numeric_text": ref_types.FeatureType.NUMERIC}
wanted_mine = {k: getattr(my_types.FeatureType, v.name) for k, v in wanted_ref.items()}
(out_ref, meta_ref), (out_mine, meta_mine) = ref_types.to_types(frame.copy(), wanted_ref), my_types.to_types(frame.copy(), wanted_mine)
assert list(out_ref.columns) == list(out_mine.columns) and np.allclose(out_ref.to_numpy(dtype=float), out_mine.to_numpy(dtype=float))
assert meta_ref.keys() == meta_mine.keys()
agreed += 1
print("configs agree:", agreed)
