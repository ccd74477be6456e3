"""RxRx1 preprocessing (parity: ``fl4health/datasets/rxrx1/preprocess.py``): keep the 50 most frequent siRNA classes,
make one client per cell type (RPE, HUVEC, HEPG2, U2OS), and store each 3-channel image as a tensor.

Storage differs from the reference on purpose: instead of one pickle per image (tens of thousands of tiny files that
each cost a Python unpickle at load time) every (client, split) is ONE ``uint8`` tensor file
``clients/{split}_data_{client}.pt`` of shape ``[N, 3, H, W]``; per-image pickles written by the reference's script are
still readable by ``load_data.py``.
"""

from __future__ import annotations

import argparse
import os
from collections.abc import Hashable
from pathlib import Path
from typing import Any

import pandas as pd
import torch

CELL_TYPES = ["RPE", "HUVEC", "HEPG2", "U2OS"]
N_TOP_SIRNA = 50
N_CHANNELS = 3  # RxRx1 has 6 fluorescence channels; following WILDS only the first three are used


def filter_and_save_data(metadata: pd.DataFrame, top_sirna_ids: list[int], cell_type: str, output_path: Path) -> None:
    keep = metadata[metadata["sirna_id"].isin(top_sirna_ids) & (metadata["cell_type"] == cell_type)]
    keep.to_csv(output_path, index=False)


def save_to_pkl(data: torch.Tensor, output_path: str) -> None:
    """Pickle one tensor (the reference's per-image storage, ``rxrx1/preprocess.py:74-83``; kept for tools that still
    write that layout — this package stores one tensor file per (client, split) instead)."""
    import pickle

    with open(output_path, "wb") as handle:
        pickle.dump(data, handle)


def load_image(row: dict[Hashable, Any], root: Path) -> torch.Tensor:
    """``[3, H, W]`` float tensor in [0, 1] assembled from the per-channel PNGs of one (experiment, plate, well, site)."""
    import numpy as np
    from PIL import Image

    channels = []
    for channel in range(1, N_CHANNELS + 1):
        path = Path(root) / "images" / str(row["experiment"]) / f"Plate{row['plate']}" / f"{row['well']}_s{row['site']}_w{channel}.png"
        if not path.exists():
            raise FileNotFoundError(f"Image not found at {path}")
        channels.append(torch.from_numpy(np.asarray(Image.open(path).convert("L"), dtype=np.uint8).copy()))
    return torch.stack(channels).float() / 255.0


def process_data(metadata: pd.DataFrame, input_dir: Path, output_dir: Path, client_num: int, type_data: str) -> None:
    images = [(load_image(row.to_dict(), Path(input_dir)) * 255.0).round().to(torch.uint8) for _, row in metadata.iterrows()]
    if images:
        torch.save(torch.stack(images), os.path.join(output_dir, f"{type_data}_data_{client_num + 1}.pt"))


def main(dataset_dir: Path) -> None:
    dataset_dir = Path(dataset_dir)
    output_dir = dataset_dir / "clients"
    output_dir.mkdir(exist_ok=True)
    data = pd.read_csv(dataset_dir / "metadata.csv")
    top = data["sirna_id"].value_counts().head(N_TOP_SIRNA).index.tolist()
    for index, cell_type in enumerate(CELL_TYPES):
        meta_path = output_dir / f"meta_data_{index + 1}.csv"
        filter_and_save_data(data, top, cell_type, meta_path)
        metadata = pd.read_csv(meta_path)
        for split in ("train", "test"):
            process_data(metadata[metadata["dataset"] == split], dataset_dir, output_dir, index, split)


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="Filter RxRx1 by the most frequent sirna_id and split clients by cell_type.")
    parser.add_argument("dataset_dir", type=str, help="Path to the dataset directory containing metadata.csv")
    main(Path(parser.parse_args().dataset_dir))
