"""RxRx1 and skin-lesion dataset pipelines on tiny synthetic copies of the on-disk layouts."""

import json
from pathlib import Path

import numpy as np
import pandas as pd
import pytest
import torch
from PIL import Image

from fl4health_b200.datasets.rxrx1 import preprocess as rx_pre
from fl4health_b200.datasets.rxrx1.load_data import create_splits, load_rxrx1_data, load_rxrx1_test_data
from fl4health_b200.datasets.skin_cancer.load_data import load_skin_cancer_data
from fl4health_b200.datasets.skin_cancer.preprocess_skin import OFFICIAL_COLUMNS, preprocess_derm7pt, preprocess_ham10000
from fl4health_b200.utils.dataset import TensorDataset


def test_rxrx1_preprocess_and_load(tmp_path: Path) -> None:
    rng = np.random.default_rng(0)
    rows = []
    for i in range(24):
        cell = rx_pre.CELL_TYPES[i % 4]
        rows.append({"experiment": f"{cell}-01", "plate": 1, "well": f"B{i:02d}", "site": 1, "sirna_id": 100 + i % 3,
                     "cell_type": cell, "dataset": "train" if i % 6 else "test"})
        folder = tmp_path / "images" / f"{cell}-01" / "Plate1"
        folder.mkdir(parents=True, exist_ok=True)
        for channel in (1, 2, 3):
            Image.fromarray(rng.integers(0, 256, (8, 8), dtype=np.uint8)).save(folder / f"B{i:02d}_s1_w{channel}.png")
    pd.DataFrame(rows).to_csv(tmp_path / "metadata.csv", index=False)
    rx_pre.main(tmp_path)
    assert (tmp_path / "clients" / "train_data_1.pt").exists()
    train, val, counts = load_rxrx1_data(tmp_path, client_num=0, batch_size=2, seed=1, placement="host")
    assert counts["train_set"] + counts["validation_set"] == 4  # 6 RPE rows, two of them are test rows
    x, y = next(iter(train))
    assert x.shape[1:] == (3, 8, 8) and 0.0 <= float(x.min()) and float(x.max()) <= 1.0 and y.dtype == torch.long
    test, test_counts = load_rxrx1_test_data(tmp_path, client_num=0, batch_size=2, placement="host")
    assert test_counts == {"eval_set": 2}
    ds = TensorDataset(torch.zeros(20, 1), torch.arange(20) % 2)
    tr, va = create_splits(ds, seed=3, train_fraction=0.8)
    assert len(tr) == 16 and len(va) == 4 and set(tr).isdisjoint(va)


def test_skin_cancer_preprocess_and_load(tmp_path: Path) -> None:
    ham = tmp_path / "HAM10000"
    ham.mkdir()
    rows = [{"image_id": f"img{i}", "dx": ["mel", "nv", "bcc", "akiec"][i % 4], "dataset": "rosendahl" if i % 2 else "vidir_modern"}
            for i in range(20)]
    pd.DataFrame(rows).to_csv(ham / "HAM10000_metadata", index=False)
    preprocess_ham10000(str(tmp_path), OFFICIAL_COLUMNS)
    payload = json.loads((ham / "HAM_rosendahl.json").read_text())
    assert payload["columns"] == OFFICIAL_COLUMNS and len(payload["data"]) == 10
    first = payload["data"][0]
    assert sum(first["extended_labels"]) == 1 and len(first["origin_labels"]) == 7
    # point the records at real (tiny) images so the loader can decode them
    rng = np.random.default_rng(0)
    for record in payload["data"]:
        path = tmp_path / Path(record["img_path"]).name
        Image.fromarray(rng.integers(0, 256, (12, 12, 3), dtype=np.uint8)).save(path)
        record["img_path"] = str(path)
    (ham / "HAM_rosendahl.json").write_text(json.dumps(payload))
    train, val, test, counts = load_skin_cancer_data(tmp_path, "Rosendahl", batch_size=2, seed=0, placement="host",
                                                     split_percents=(0.6, 0.2, 0.2))
    assert counts == {"train_set": 6, "validation_set": 2, "test_set": 2}
    x, y = next(iter(val))
    assert x.shape == (2, 3, 256, 256) and int(y.max()) < 8
    with pytest.raises(ValueError):
        load_skin_cancer_data(tmp_path, "Nowhere", 2)
    derm = tmp_path / "Derm7pt" / "meta"
    derm.mkdir(parents=True)
    pd.DataFrame({"derm": ["a.jpg", "b.jpg", "c.jpg"], "diagnosis": ["melanoma", "miscellaneous", "clark nevus"]}).to_csv(
        derm / "meta_core.csv", index=False)
    preprocess_derm7pt(str(tmp_path), OFFICIAL_COLUMNS)
    derm_payload = json.loads((tmp_path / "Derm7pt" / "Derm7pt.json").read_text())
    assert len(derm_payload["data"]) == 2  # the "miscellaneous" lesion has no class in the shared label space
