"""Skin-lesion client loaders (parity: ``fl4health/datasets/skin_cancer/load_data.py:26-188``)."""

from __future__ import annotations

import json
import random
from collections.abc import Callable
from concurrent.futures import ThreadPoolExecutor
from logging import INFO
from pathlib import Path
from typing import Any

import torch

from fl4health_b200.common.logger import log
from fl4health_b200.engine.data import BatchedTensorLoader
from fl4health_b200.utils.dataset import TensorDataset
from fl4health_b200.utils.dataset_converter import DatasetConverter
from fl4health_b200.utils.sampler import LabelBasedSampler

DATASET_FILES = {
    "Barcelona": ("ISIC_2019", "ISIC_19_Barcelona.json"),
    "Rosendahl": ("HAM10000", "HAM_rosendahl.json"),
    "Vienna": ("HAM10000", "HAM_vienna.json"),
    "UFES": ("PAD-UFES-20", "PAD_UFES_20.json"),
    "Canada": ("Derm7pt", "Derm7pt.json"),
}
IMAGE_SIZE = [256, 256]


def _default_transforms() -> tuple[Callable, Callable]:
    from torchvision import transforms

    tail = [transforms.Resize(IMAGE_SIZE), transforms.ToTensor(), transforms.Normalize((0.0, 0.0, 0.0), (1.0, 1.0, 1.0))]
    train = transforms.Compose([
        transforms.RandomHorizontalFlip(), transforms.RandomVerticalFlip(), transforms.RandomRotation(20),
        transforms.ColorJitter(brightness=32.0 / 255.0, saturation=0.5), *tail,
    ])
    return train, transforms.Compose(tail)


def load_image(item: dict[str, Any], transform: Callable | None) -> tuple[torch.Tensor, int]:
    from PIL import Image
    from torchvision import transforms

    image = Image.open(item["img_path"]).convert("RGB")
    tensor = transform(image) if transform else transforms.ToTensor()(image)
    assert isinstance(tensor, torch.Tensor), f"Image at {item['img_path']} is not a Tensor"
    return tensor, int(torch.tensor(item["extended_labels"]).argmax().item())


def construct_skin_cancer_tensor_dataset(data: list[dict[str, Any]], transform: Callable | None = None, num_workers: int = 8) -> TensorDataset:
    """Decode + transform every image once (thread pool), keep the result as one tensor."""
    with ThreadPoolExecutor(max_workers=num_workers) as pool:
        loaded = list(pool.map(lambda item: load_image(item, transform), data))
    return TensorDataset(torch.stack([d for d, _ in loaded]), torch.tensor([t for _, t in loaded]))


def load_skin_cancer_data(
    data_dir: Path, dataset_name: str, batch_size: int, split_percents: tuple[float, float, float] = (0.7, 0.15, 0.15),
    sampler: LabelBasedSampler | None = None, train_transform: Callable | None = None, val_transform: Callable | None = None,
    test_transform: Callable | None = None, dataset_converter: DatasetConverter | None = None, seed: int | None = None,
    placement: str = "pinned", device: torch.device | str | None = None,
) -> tuple[BatchedTensorLoader, BatchedTensorLoader, BatchedTensorLoader, dict[str, int]]:
    if sum(split_percents) != 1.0:
        raise ValueError("The split percentages must sum to 1.0")
    if dataset_name not in DATASET_FILES:
        raise ValueError(f"Dataset {dataset_name} not found in available datasets.")
    dataset_path = Path(data_dir).joinpath(*DATASET_FILES[dataset_name])
    if not dataset_path.exists():
        raise FileNotFoundError(f"Dataset file {dataset_path} does not exist. Run datasets/skin_cancer/preprocess_skin.py first.")
    log(INFO, f"Data directory: {dataset_path!s}")
    with open(dataset_path) as handle:
        records = json.load(handle)["data"]
    random.Random(seed).shuffle(records) if seed is not None else random.shuffle(records)
    n_train, n_val = int(split_percents[0] * len(records)), int(split_percents[1] * len(records))
    parts = records[:n_train], records[n_train:n_train + n_val], records[n_train + n_val:]
    default_train, default_eval = _default_transforms()
    chosen = (train_transform or default_train, val_transform or default_eval, test_transform or default_eval)
    datasets = [construct_skin_cancer_tensor_dataset(part, transform=t) for part, t in zip(parts, chosen)]
    if sampler is not None:
        datasets = [sampler.subsample(ds) for ds in datasets]
    if dataset_converter is not None:
        import copy

        datasets = [copy.copy(dataset_converter).convert_dataset(ds) for ds in datasets]
    train_ds, val_ds, test_ds = datasets
    loaders = (
        BatchedTensorLoader(train_ds, batch_size, shuffle=True, placement=placement, device=device),
        BatchedTensorLoader(val_ds, batch_size, placement=placement, device=device),
        BatchedTensorLoader(test_ds, batch_size, placement=placement, device=device),
    )
    return (*loaders, {"train_set": len(train_ds), "validation_set": len(val_ds), "test_set": len(test_ds)})
