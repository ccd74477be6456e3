"""Medical Segmentation Decathlon task registry (parity: ``fl4health/utils/msd_dataset_sources.py``): archive URL,
md5 and number of segmentation classes (background included) per task, built from one table."""

from __future__ import annotations

from enum import Enum

_BUCKET = "https://msd-for-monai.s3-us-west-2.amazonaws.com"

# (task folder, md5 of the .tar, number of labels incl. background)
_TASKS = (
    ("Task01_BrainTumour", "240a19d752f0d9e9101544901065d872", 4),
    ("Task02_Heart", "06ee59366e1e5124267b774dbd654057", 2),
    ("Task03_Liver", "a90ec6c4aa7f6a3d087205e23d4e6397", 3),
    ("Task04_Hippocampus", "9d24dba78a72977dbd1d2e110310f31b", 3),
    ("Task05_Prostate", "35138f08b1efaef89d7424d2bcc928db", 3),
    ("Task06_Lung", "8afd997733c7fc0432f71255ba4e52dc", 2),
    ("Task07_Pancreas", "4f7080cfca169fa8066d17ce6eb061e4", 3),
    ("Task08_HepaticVessel", "641d79e80ec66453921d997fbf12a29c", 3),
    ("Task09_Spleen", "410d4a301da4e5b2f6f86ec3ddba524e", 2),
    ("Task10_Colon", "bad7a188931dc2f6acf72b08eb6202d0", 2),
)

MsdDataset = Enum("MsdDataset", {name.upper(): name for name, _, _ in _TASKS})  # type: ignore[misc]


def get_msd_dataset_enum(dataset_name: str) -> MsdDataset:
    return MsdDataset(dataset_name)


msd_urls = {MsdDataset(name): f"{_BUCKET}/{name}.tar" for name, _, _ in _TASKS}
msd_md5_hashes = {MsdDataset(name): md5 for name, md5, _ in _TASKS}
msd_num_labels = {MsdDataset(name): labels for name, _, labels in _TASKS}
