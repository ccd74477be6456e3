"""Skin-lesion federation preprocessing (parity: ``fl4health/datasets/skin_cancer/preprocess_skin.py``; the client
split follows Yang et al., arXiv:2207.03075): ISIC-2019 (Barcelona), HAM10000 (Rosendahl / Vienna), PAD-UFES-20 and
Derm7pt are mapped onto one 8-class label space and written as ``{columns, original_columns, data: [{img_path,
origin_labels, extended_labels}]}`` JSON files.  The four sources are described by ONE table (``SOURCES``) instead of
four near-identical functions."""

from __future__ import annotations

import json
import os
from collections.abc import Callable
from dataclasses import dataclass, field
from typing import Any

import pandas as pd

OFFICIAL_COLUMNS = ["MEL", "NV", "BCC", "AK", "BKL", "DF", "VASC", "SCC"]


def save_to_json(data: dict[str, Any], path: str) -> None:
    with open(path, "w", encoding="utf-8") as handle:
        json.dump(data, handle, indent="\t")


def process_client_data(
    dataframe: pd.DataFrame, client_name: str, data_path: str, image_path_func: Callable[[pd.Series], str],
    label_map_func: Callable[[pd.Series], str], original_columns: list[str], official_columns: list[str],
) -> None:
    records = []
    for _, row in dataframe.iterrows():
        label = label_map_func(row)
        if label not in official_columns:  # e.g. Derm7pt "miscellaneous": no counterpart in the shared label space
            continue
        records.append({
            "img_path": image_path_func(row),
            "origin_labels": [int(c == label) for c in original_columns],
            "extended_labels": [int(c == label) for c in official_columns],
        })
    save_to_json({"columns": official_columns, "original_columns": original_columns, "data": records},
                 os.path.join(data_path, f"{client_name}.json"))


@dataclass
class _Source:
    folder: str
    metadata: str                      # csv path relative to the folder
    image_column: str
    image_prefix: tuple[str, ...]      # path components prepended to the image id
    image_suffix: str
    label_column: str
    label_map: dict[str, str]
    original_columns: list[str]
    clients: dict[str, Callable[[pd.DataFrame], pd.DataFrame]] = field(default_factory=dict)  # client name -> row filter


_ROOT = ("fl4health", "datasets", "skin_cancer")
_NEVI = ("blue nevus", "clark nevus", "combined nevus", "congenital nevus", "dermal nevus", "recurrent nevus", "reed or spitz nevus")
_MELANOMAS = ("melanoma", "melanoma (0.76 to 1.5 mm)", "melanoma (in situ)", "melanoma (less than 0.76 mm)",
              "melanoma (more than 1.5 mm)", "melanoma metastasis")

SOURCES: dict[str, _Source] = {
    "HAM10000": _Source(
        "HAM10000", "HAM10000_metadata", "image_id", (*_ROOT, "HAM10000"), ".jpg", "dx",
        {"akiec": "AK", "bcc": "BCC", "bkl": "BKL", "df": "DF", "mel": "MEL", "nv": "NV", "vasc": "VASC"},
        ["MEL", "NV", "BCC", "AK", "BKL", "DF", "VASC"],
        {"HAM_rosendahl": lambda df: df[df["dataset"] == "rosendahl"], "HAM_vienna": lambda df: df[df["dataset"] != "rosendahl"]},
    ),
    "PAD-UFES-20": _Source(
        "PAD-UFES-20", "metadata.csv", "img_id", (*_ROOT, "PAD-UFES-20"), "", "diagnostic",
        {"ACK": "AK", "BCC": "BCC", "MEL": "MEL", "NEV": "NV", "SCC": "SCC", "SEK": "BKL"},
        ["MEL", "NV", "BCC", "AK", "BKL", "SCC"], {"PAD_UFES_20": lambda df: df},
    ),
    "Derm7pt": _Source(
        "Derm7pt", os.path.join("meta", "meta_core.csv"), "derm", (*_ROOT, "Derm7pt", "images"), "", "diagnosis",
        {"basal cell carcinoma": "BCC", "dermatofibroma": "DF", "seborrheic keratosis": "BKL", "vascular lesion": "VASC",
         "lentigo": "MISC", "melanosis": "MISC", "miscellaneous": "MISC",
         **dict.fromkeys(_NEVI, "NV"), **dict.fromkeys(_MELANOMAS, "MEL")},
        ["MEL", "NV", "BCC", "BKL", "DF", "VASC"], {"Derm7pt": lambda df: df},
    ),
}


def _image_path(source: _Source, row: pd.Series) -> str:
    return os.path.join(*source.image_prefix, str(row[source.image_column]) + source.image_suffix)


def _label(source: _Source, row: pd.Series) -> str:
    return source.label_map[row[source.label_column]]


# Public per-dataset helpers with the reference's names (preprocess_skin.py:120-152, 194-225, 253-300): views over SOURCES.
def ham_image_path_func(row: pd.Series) -> str:
    return _image_path(SOURCES["HAM10000"], row)


def ham_label_map_func(row: pd.Series) -> str:
    return _label(SOURCES["HAM10000"], row)


def pad_image_path_func(row: pd.Series) -> str:
    return _image_path(SOURCES["PAD-UFES-20"], row)


def pad_label_map_func(row: pd.Series) -> str:
    return _label(SOURCES["PAD-UFES-20"], row)


def derm7pt_image_path_func(row: pd.Series) -> str:
    return _image_path(SOURCES["Derm7pt"], row)


def derm7pt_label_map_func(row: pd.Series) -> str:
    return _label(SOURCES["Derm7pt"], row)


def _preprocess_source(data_path: str, source: _Source, official_columns: list[str]) -> None:
    folder = os.path.join(data_path, source.folder)
    frame = pd.read_csv(os.path.join(folder, source.metadata))
    for client_name, row_filter in source.clients.items():
        process_client_data(
            row_filter(frame).reset_index(drop=True), client_name, folder,
            lambda row, s=source: _image_path(s, row), lambda row, s=source: _label(s, row), source.original_columns,
            official_columns,
        )


def preprocess_ham10000(data_path: str, official_columns: list[str]) -> None:
    _preprocess_source(data_path, SOURCES["HAM10000"], official_columns)


def preprocess_pad_ufes_20(data_path: str, official_columns: list[str]) -> None:
    _preprocess_source(data_path, SOURCES["PAD-UFES-20"], official_columns)


def preprocess_derm7pt(data_path: str, official_columns: list[str]) -> None:
    _preprocess_source(data_path, SOURCES["Derm7pt"], official_columns)


def preprocess_isic_2019(data_path: str, official_columns: list[str]) -> None:
    """ISIC-2019 ships one-hot ground truth already; only the Barcelona (BCN) lesions form the client."""
    folder = os.path.join(data_path, "ISIC_2019")
    truth = pd.read_csv(os.path.join(folder, "ISIC_2019_Training_GroundTruth.csv"))
    meta = pd.read_csv(os.path.join(folder, "ISIC_2019_Training_Metadata.csv"))
    barcelona_images = meta[meta["lesion_id"].fillna("").str.contains("BCN")]["image"]
    core = truth[truth["image"].isin(barcelona_images)].reset_index(drop=True)
    core.to_csv(os.path.join(folder, "ISIC_2019_core.csv"), mode="w")
    image_dir = os.path.join(data_path, "ISIC_2019", "ISIC_2019_Training_Input")
    records = []
    for _, row in core.iterrows():
        labels = [row[c].item() if hasattr(row[c], "item") else row[c] for c in official_columns]
        records.append({"img_path": os.path.join(image_dir, f"{row['image']}.jpg"), "origin_labels": labels, "extended_labels": labels})
    save_to_json({"columns": official_columns, "original_columns": official_columns, "data": records},
                 os.path.join(folder, "ISIC_19_Barcelona.json"))


if __name__ == "__main__":
    root = os.path.join(*_ROOT)
    preprocess_isic_2019(root, OFFICIAL_COLUMNS)
    preprocess_ham10000(root, OFFICIAL_COLUMNS)
    preprocess_pad_ufes_20(root, OFFICIAL_COLUMNS)
    preprocess_derm7pt(root, OFFICIAL_COLUMNS)
