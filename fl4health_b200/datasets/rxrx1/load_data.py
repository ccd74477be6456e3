"""RxRx1 client loaders (parity: ``fl4health/datasets/rxrx1/load_data.py:18-196``): stratified train/val split of a
client's training images, test loader, label-frequency logging."""

from __future__ import annotations

import os
import pickle
from collections.abc import Callable
from logging import INFO
from pathlib import Path

import numpy as np
import pandas as pd
import torch

from fl4health_b200.common.logger import log
from fl4health_b200.engine.data import BatchedTensorLoader
from fl4health_b200.utils.dataset import TensorDataset, select_by_indices


def _load_images(data_path: Path, client_num: int, dataset_type: str, count: int) -> torch.Tensor:
    packed = Path(data_path) / "clients" / f"{dataset_type}_data_{client_num + 1}.pt"
    if packed.exists():  # one tensor per (client, split): see preprocess.py
        images = torch.load(packed)
        return images.float() / 255.0 if images.dtype == torch.uint8 else images.float()
    folder = Path(data_path) / "clients" / f"{dataset_type}_data_{client_num + 1}"  # the reference's per-image pickles
    tensors = []
    for index in range(count):
        with open(os.path.join(folder, f"image_{index}.pkl"), "rb") as handle:
            tensors.append(torch.as_tensor(pickle.load(handle)).float())
    return torch.stack(tensors)


def construct_rxrx1_tensor_dataset(
    metadata: pd.DataFrame, data_path: Path, client_num: int, dataset_type: str, transform: Callable | None = None
) -> tuple[TensorDataset, dict[int, int]]:
    """Labels are re-indexed to 0..C-1 over the siRNA ids present in ``metadata``; returns the inverse map too."""
    label_map = {label: index for index, label in enumerate(sorted(metadata["sirna_id"].unique()))}
    original_label_map = {new: original for original, new in label_map.items()}
    subset = metadata[metadata["dataset"] == dataset_type]
    targets = torch.tensor(subset["sirna_id"].map(label_map).tolist(), dtype=torch.long)
    data = _load_images(data_path, client_num, dataset_type, len(targets))
    return TensorDataset(data, targets, transform), original_label_map


def label_frequency(dataset: TensorDataset, original_label_map: dict[int, int]) -> None:
    assert isinstance(dataset, TensorDataset) and isinstance(dataset.targets, torch.Tensor), "Dataset must be a TensorDataset"
    labels, counts = torch.unique(dataset.targets, return_counts=True)
    for label, count in zip(labels.tolist(), counts.tolist()):
        log(INFO, f"Label {label} (original: {original_label_map.get(label)}): {count} samples")


def create_splits(dataset: TensorDataset, seed: int | None = None, train_fraction: float = 0.8) -> tuple[list[int], list[int]]:
    """Stratified split: within every label a ``train_fraction`` share of the (shuffled) indices goes to train."""
    assert isinstance(dataset.targets, torch.Tensor)
    train_indices: list[int] = []
    val_indices: list[int] = []
    for label in torch.unique(dataset.targets).tolist():
        indices = (dataset.targets == label).nonzero().reshape(-1).tolist()
        (np.random.default_rng(seed) if seed is not None else np.random).shuffle(indices)
        split = int(len(indices) * train_fraction)
        train_indices.extend(indices[:split])
        val_indices.extend(indices[split:])
    if not val_indices:
        log(INFO, "Warning: Validation set is empty. Consider changing the train_fraction parameter.")
    return train_indices, val_indices


def load_rxrx1_data(
    data_path: Path, client_num: int, batch_size: int, seed: int | None = None, train_val_split: float = 0.8,
    num_workers: int = 0, placement: str = "pinned", device: torch.device | str | None = None,  # noqa: ARG001
) -> tuple[BatchedTensorLoader, BatchedTensorLoader, dict[str, int]]:
    metadata = pd.read_csv(f"{data_path}/clients/meta_data_{client_num + 1}.csv")
    dataset, _ = construct_rxrx1_tensor_dataset(metadata, Path(data_path), client_num, "train")
    train_indices, val_indices = create_splits(dataset, seed=seed, train_fraction=train_val_split)
    train_set = select_by_indices(dataset, torch.tensor(train_indices, dtype=torch.long))
    val_set = select_by_indices(dataset, torch.tensor(val_indices, dtype=torch.long))
    train_loader = BatchedTensorLoader(train_set, batch_size, shuffle=True, placement=placement, device=device)
    val_loader = BatchedTensorLoader(val_set, batch_size, placement=placement, device=device)
    return train_loader, val_loader, {"train_set": len(train_set), "validation_set": len(val_set)}


def load_rxrx1_test_data(
    data_path: Path, client_num: int, batch_size: int, num_workers: int = 0, placement: str = "pinned",  # noqa: ARG001
    device: torch.device | str | None = None,
) -> tuple[BatchedTensorLoader, dict[str, int]]:
    metadata = pd.read_csv(f"{data_path}/clients/meta_data_{client_num + 1}.csv")
    dataset, _ = construct_rxrx1_tensor_dataset(metadata, Path(data_path), client_num, "test")
    return BatchedTensorLoader(dataset, batch_size, placement=placement, device=device), {"eval_set": len(dataset)}
