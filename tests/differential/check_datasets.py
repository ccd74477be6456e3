"""Domain dataset pipelines vs the reference on tiny synthetic copies of the raw on-disk layouts: the four skin-lesion
preprocessors (ISIC-2019 Barcelona, HAM10000 per source, PAD-UFES-20, Derm7pt -> the shared 8-class JSON records) and the
RxRx1 helpers (per-client filtering of the metadata, the seeded train / validation split)."""
import json
import shutil
import tempfile
from pathlib import Path

import numpy as np
import pandas as pd
import torch

import fl4health.datasets.rxrx1.load_data as ref_rx_load
import fl4health.datasets.rxrx1.preprocess as ref_rx_pre
import fl4health.datasets.skin_cancer.preprocess_skin as ref_skin
import fl4health.utils.dataset as ref_ds
import fl4health_b200.datasets.rxrx1.load_data as my_rx_load
import fl4health_b200.datasets.rxrx1.preprocess as my_rx_pre
import fl4health_b200.datasets.skin_cancer.preprocess_skin as my_skin
import fl4health_b200.utils.dataset as my_ds

agreed = 0
COLUMNS = ["MEL", "NV", "BCC", "AK", "BKL", "DF", "VASC", "SCC"]
rng = np.random.default_rng(61)


def raw_layout(root: Path) -> None:
    isic = root / "ISIC_2019"
    isic.mkdir(parents=True)
    images = [f"ISIC_{i:07d}" for i in range(24)]
    labels = rng.integers(0, len(COLUMNS) + 1, 24)
    truth = pd.DataFrame({"image": images, **{c: (labels == j).astype(float) for j, c in enumerate(COLUMNS + ["UNK"])}})
    truth.to_csv(isic / "ISIC_2019_Training_GroundTruth.csv", index=False)
    pd.DataFrame({"image": images, "lesion_id": [f"BCN_{i:04d}" if i % 3 else (f"HAM_{i:04d}" if i % 2 else None) for i in range(24)]}).to_csv(
        isic / "ISIC_2019_Training_Metadata.csv", index=False)
    ham = root / "HAM10000"
    ham.mkdir()
    pd.DataFrame([{"image_id": f"img{i}", "dx": ["mel", "nv", "bcc", "akiec", "bkl", "df", "vasc"][i % 7],
                   "dataset": ["rosendahl", "vidir_modern", "vidir_molemax", "vienna_dias"][i % 4]} for i in range(40)]).to_csv(ham / "HAM10000_metadata", index=False)
    pad = root / "PAD-UFES-20"
    pad.mkdir()
    pd.DataFrame({"img_id": [f"PAT_{i}.png" for i in range(18)], "diagnostic": [["ACK", "BCC", "MEL", "NEV", "SCC", "SEK"][i % 6] for i in range(18)]}).to_csv(pad / "metadata.csv", index=False)
    derm = root / "Derm7pt" / "meta"
    derm.mkdir(parents=True)
    # (no lentigo / melanosis / miscellaneous rows: the reference maps them to "MISC", which is not one of its columns, and
    # raises; ours skips such lesions -- tests/test_domain_datasets.py)
    diagnoses = ["basal cell carcinoma", "blue nevus", "clark nevus", "dermatofibroma", "melanoma", "melanoma (in situ)",
                 "reed or spitz nevus", "seborrheic keratosis", "vascular lesion", "melanoma metastasis"]
    pd.DataFrame({"derm": [f"d{i}.jpg" for i in range(len(diagnoses))], "diagnosis": diagnoses}).to_csv(derm / "meta_core.csv", index=False)


roots = {}
for label, module in (("reference", ref_skin), ("ours", my_skin)):
    root = Path(tempfile.mkdtemp(prefix=f"skin_{label}_")) / "data"
    if label == "reference":
        raw_layout(root)
    else:
        shutil.copytree(roots["reference"], root)  # the same raw files (before either side has written its outputs)
        for produced in root.rglob("*.json"):
            produced.unlink()
    roots[label] = root
    if label == "reference":
        pristine = Path(tempfile.mkdtemp(prefix="skin_raw_")) / "data"
        shutil.copytree(root, pristine)
    module.preprocess_isic_2019(str(root), COLUMNS)
    module.preprocess_ham10000(str(root), COLUMNS)
    module.preprocess_pad_ufes_20(str(root), COLUMNS)
    module.preprocess_derm7pt(str(root), COLUMNS)

produced_ref = sorted(p.relative_to(roots["reference"]) for p in roots["reference"].rglob("*.json"))
produced_mine = sorted(p.relative_to(roots["ours"]) for p in roots["ours"].rglob("*.json"))
assert produced_ref == produced_mine and len(produced_ref) == 5, (produced_ref, produced_mine)
for relative in produced_ref:
    theirs, ours = json.loads((roots["reference"] / relative).read_text()), json.loads((roots["ours"] / relative).read_text())
    assert theirs["columns"] == ours["columns"] and theirs["original_columns"] == ours["original_columns"], relative
    assert len(theirs["data"]) == len(ours["data"]), (relative, len(theirs["data"]), len(ours["data"]))
    for a, b in zip(theirs["data"], ours["data"]):
        # image paths are rooted differently (each side ran on its own copy): compare what follows the dataset folder
        assert Path(a["img_path"]).name == Path(b["img_path"]).name, (relative, a["img_path"], b["img_path"])
        assert [float(v) for v in a["origin_labels"]] == [float(v) for v in b["origin_labels"]], relative
        assert [float(v) for v in a["extended_labels"]] == [float(v) for v in b["extended_labels"]], relative
    agreed += 1

# RxRx1: which rows each client keeps, and the seeded split
metadata = pd.DataFrame({"experiment": [f"{cell}-0{1 + i % 2}" for i, cell in enumerate(["HEPG2", "HUVEC", "RPE", "U2OS"] * 15)],
                         "cell_type": ["HEPG2", "HUVEC", "RPE", "U2OS"] * 15, "sirna_id": list(rng.integers(0, 9, 60)), "well": [f"B{i:02d}" for i in range(60)],
                         "plate": 1, "site": 1, "dataset": ["train", "test"] * 30})
top = [0, 1, 2, 3]
for cell in ("RPE", "HUVEC"):
    out_ref, out_mine = Path(tempfile.mkdtemp()) / "ref.csv", Path(tempfile.mkdtemp()) / "mine.csv"
    ref_rx_pre.filter_and_save_data(metadata.copy(), top, cell, out_ref)
    my_rx_pre.filter_and_save_data(metadata.copy(), top, cell, out_mine)
    a, b = pd.read_csv(out_ref), pd.read_csv(out_mine)
    assert list(a.columns) == list(b.columns) and a.equals(b), (cell, a.shape, b.shape)
    agreed += 1
for seed, fraction in ((3, 0.8), (11, 0.5)):
    labels = torch.arange(60) % 4
    a_train, a_val = ref_rx_load.create_splits(ref_ds.TensorDataset(torch.zeros(60, 1), labels), seed=seed, train_fraction=fraction)
    b_train, b_val = my_rx_load.create_splits(my_ds.TensorDataset(torch.zeros(60, 1), labels), seed=seed, train_fraction=fraction)
    assert sorted(a_train) == sorted(b_train) and sorted(a_val) == sorted(b_val), (seed, fraction)
    agreed += 1
print("configs agree:", agreed)
