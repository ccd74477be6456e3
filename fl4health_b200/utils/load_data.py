"""MNIST / CIFAR-10 loaders (parity: ``fl4health/utils/load_data.py:26-300``).

Differences by design:

* the raw files are parsed directly (IDX for MNIST, pickled batches for CIFAR-10) from torchvision's on-disk layout
  (``<data_dir>/MNIST/raw/*-ubyte[.gz]``, ``<data_dir>/cifar-10-batches-py/*``); if they are absent and torchvision
  can download them it is asked to, otherwise a ``FileNotFoundError`` explains what is expected (no silent download
  attempts on air-gapped nodes);
* the default ``ToTensor + Normalize(0.5, 0.5)`` pipeline is applied ONCE to the whole tensor (a vectorised
  ``batch_transform``) instead of per sample in ``__getitem__``;
* loaders are ``BatchedTensorLoader``s (pinned / device-resident datasets, see ``engine/data.py``).
"""

from __future__ import annotations

import gzip
import pickle
import random
import struct
from collections.abc import Callable
from logging import INFO
from pathlib import Path

import numpy as np
import torch

from fl4health_b200.common.logger import log
from fl4health_b200.engine.data import BatchedTensorLoader
from fl4health_b200.utils.dataset import TensorDataset
from fl4health_b200.utils.dataset_converter import DatasetConverter
from fl4health_b200.utils.sampler import LabelBasedSampler


class ToNumpy:
    def __call__(self, tensor: torch.Tensor) -> np.ndarray:
        return tensor.numpy()


def split_data_and_targets(
    data: torch.Tensor, targets: torch.Tensor, validation_proportion: float = 0.2, hash_key: int | None = None
) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Random train/validation split (deterministic for a given ``hash_key``)."""
    total = data.shape[0]
    train_size = int(total * (1 - validation_proportion))
    rng = random.Random(hash_key) if hash_key is not None else random
    train_indices = rng.sample(range(total), train_size)
    mask = torch.ones(total, dtype=torch.bool)
    mask[train_indices] = False
    val_indices = mask.nonzero().reshape(-1)
    train_index = torch.tensor(train_indices, dtype=torch.int64)
    return data[train_index], targets[train_index], data[val_indices], targets[val_indices]


# ---------------------------------------------------------------------------------------------------------------
# raw readers
# ---------------------------------------------------------------------------------------------------------------
def _read_idx(path: Path) -> np.ndarray:
    opener = gzip.open if path.suffix == ".gz" else open
    with opener(path, "rb") as handle:
        _, _, dtype_code, ndim = struct.unpack(">BBBB", handle.read(4))
        assert dtype_code == 0x08, "only unsigned-byte IDX files are supported"
        shape = struct.unpack(">" + "I" * ndim, handle.read(4 * ndim))
        return np.frombuffer(handle.read(), dtype=np.uint8).reshape(shape)


def _find(data_dir: Path, candidates: list[str]) -> Path | None:
    for name in candidates:
        for root in (data_dir, data_dir / "MNIST" / "raw", data_dir / "raw"):
            if (root / name).exists():
                return root / name
    return None


def get_mnist_data_and_target_tensors(data_dir: Path, train: bool) -> tuple[torch.Tensor, torch.Tensor]:
    prefix = "train" if train else "t10k"
    images = _find(data_dir, [f"{prefix}-images-idx3-ubyte", f"{prefix}-images-idx3-ubyte.gz"])
    labels = _find(data_dir, [f"{prefix}-labels-idx1-ubyte", f"{prefix}-labels-idx1-ubyte.gz"])
    if images is None or labels is None:
        try:  # let torchvision fetch it when a network is available
            from torchvision.datasets import MNIST

            dataset = MNIST(str(data_dir), train=train, download=True)
            return torch.Tensor(dataset.data), torch.Tensor(dataset.targets).long()
        except Exception as exc:  # noqa: BLE001
            raise FileNotFoundError(
                f"MNIST files not found under {data_dir} (expected {prefix}-images-idx3-ubyte[.gz] and "
                f"{prefix}-labels-idx1-ubyte[.gz], e.g. in {data_dir}/MNIST/raw) and download failed: {exc}"
            ) from exc
    return torch.from_numpy(_read_idx(images).copy()).float(), torch.from_numpy(_read_idx(labels).copy()).long()


def get_cifar10_data_and_target_tensors(data_dir: Path, train: bool) -> tuple[torch.Tensor, torch.Tensor]:
    """Returns uint8 images ``[N, 32, 32, 3]`` (HWC, as torchvision stores them) and int64 labels."""
    root = data_dir / "cifar-10-batches-py"
    names = [f"data_batch_{i}" for i in range(1, 6)] if train else ["test_batch"]
    if not all((root / name).exists() for name in names):
        try:
            from torchvision.datasets import CIFAR10

            dataset = CIFAR10(str(data_dir), train=train, download=True)
            return torch.from_numpy(dataset.data), torch.Tensor(dataset.targets).long()
        except Exception as exc:  # noqa: BLE001
            raise FileNotFoundError(f"CIFAR-10 batches not found under {root} and download failed: {exc}") from exc
    data, labels = [], []
    for name in names:
        with open(root / name, "rb") as handle:
            entry = pickle.load(handle, encoding="latin1")
        data.append(np.asarray(entry["data"], dtype=np.uint8).reshape(-1, 3, 32, 32))
        labels.extend(entry.get("labels", entry.get("fine_labels")))
    images = np.concatenate(data).transpose(0, 2, 3, 1)
    return torch.from_numpy(np.ascontiguousarray(images)), torch.tensor(labels, dtype=torch.int64)


# ---------------------------------------------------------------------------------------------------------------
# vectorised default transforms (== ToTensor() + Normalize(0.5, 0.5) of the reference, but on whole batches)
# ---------------------------------------------------------------------------------------------------------------
def mnist_batch_transform(batch: torch.Tensor) -> torch.Tensor:
    """[B, 28, 28] raw 0..255 -> [B, 1, 28, 28] in [-1, 1]."""
    return (batch.float().unsqueeze(1) / 255.0 - 0.5) / 0.5


def cifar10_batch_transform(batch: torch.Tensor) -> torch.Tensor:
    """[B, 32, 32, 3] uint8 -> [B, 3, 32, 32] in [-1, 1]."""
    return (batch.permute(0, 3, 1, 2).float() / 255.0 - 0.5) / 0.5


def _make_dataset(data: torch.Tensor, targets: torch.Tensor, transform: Callable | None, target_transform: Callable | None,
                  default_batch_transform: Callable) -> TensorDataset:
    if transform is None:
        return TensorDataset(data, targets, target_transform=target_transform, batch_transform=default_batch_transform)
    return TensorDataset(data, targets, transform=transform, target_transform=target_transform)


def get_train_and_val_mnist_datasets(
    data_dir: Path, transform: Callable | None = None, target_transform: Callable | None = None,
    validation_proportion: float = 0.2, hash_key: int | None = None,
) -> tuple[TensorDataset, TensorDataset]:
    data, targets = get_mnist_data_and_target_tensors(data_dir, True)
    train_x, train_y, val_x, val_y = split_data_and_targets(data, targets, validation_proportion, hash_key)
    return (_make_dataset(train_x, train_y, transform, target_transform, mnist_batch_transform),
            _make_dataset(val_x, val_y, transform, target_transform, mnist_batch_transform))


def get_train_and_val_cifar10_datasets(
    data_dir: Path, transform: Callable | None = None, target_transform: Callable | None = None,
    validation_proportion: float = 0.2, hash_key: int | None = None,
) -> tuple[TensorDataset, TensorDataset]:
    data, targets = get_cifar10_data_and_target_tensors(data_dir, True)
    train_x, train_y, val_x, val_y = split_data_and_targets(data, targets, validation_proportion, hash_key)
    return (_make_dataset(train_x, train_y, transform, target_transform, cifar10_batch_transform),
            _make_dataset(val_x, val_y, transform, target_transform, cifar10_batch_transform))


def _finish(
    training_set: TensorDataset, validation_set: TensorDataset | None, batch_size: int, sampler: LabelBasedSampler | None,
    dataset_converter: DatasetConverter | None, placement: str, device: torch.device | str | None,
) -> tuple[BatchedTensorLoader, BatchedTensorLoader | None, dict[str, int]]:
    if sampler is not None:
        training_set = sampler.subsample(training_set)
        validation_set = sampler.subsample(validation_set) if validation_set is not None else None
    if dataset_converter is not None:
        training_set = dataset_converter.convert_dataset(training_set)
        if validation_set is not None:
            import copy

            validation_set = copy.copy(dataset_converter).convert_dataset(validation_set)
    train_loader = BatchedTensorLoader(training_set, batch_size, shuffle=True, placement=placement, device=device)
    val_loader = BatchedTensorLoader(validation_set, batch_size, placement=placement, device=device) if validation_set is not None else None
    return train_loader, val_loader, {"train_set": len(training_set), "validation_set": len(validation_set) if validation_set is not None else 0}


def load_mnist_data(
    data_dir: Path, batch_size: int, sampler: LabelBasedSampler | None = None, transform: Callable | None = None,
    target_transform: Callable | None = None, dataset_converter: DatasetConverter | None = None,
    validation_proportion: float = 0.2, hash_key: int | None = None, placement: str = "host",
    device: torch.device | str | None = None,
) -> tuple[BatchedTensorLoader, BatchedTensorLoader, dict[str, int]]:
    log(INFO, f"Data directory: {data_dir!s}")
    train, val = get_train_and_val_mnist_datasets(data_dir, transform, target_transform, validation_proportion, hash_key)
    train_loader, val_loader, counts = _finish(train, val, batch_size, sampler, dataset_converter, placement, device)
    assert val_loader is not None
    return train_loader, val_loader, counts


def load_mnist_test_data(
    data_dir: Path, batch_size: int, sampler: LabelBasedSampler | None = None, transform: Callable | None = None,
    placement: str = "host", device: torch.device | str | None = None,
) -> tuple[BatchedTensorLoader, dict[str, int]]:
    log(INFO, f"Data directory: {data_dir!s}")
    data, targets = get_mnist_data_and_target_tensors(data_dir, False)
    test_set = _make_dataset(data, targets, transform, None, mnist_batch_transform)
    if sampler is not None:
        test_set = sampler.subsample(test_set)
    return BatchedTensorLoader(test_set, batch_size, placement=placement, device=device), {"eval_set": len(test_set)}


def load_cifar10_data(
    data_dir: Path, batch_size: int, sampler: LabelBasedSampler | None = None, validation_proportion: float = 0.2,
    hash_key: int | None = None, placement: str = "host", device: torch.device | str | None = None,
) -> tuple[BatchedTensorLoader, BatchedTensorLoader, dict[str, int]]:
    log(INFO, f"Data directory: {data_dir!s}")
    train, val = get_train_and_val_cifar10_datasets(data_dir, None, None, validation_proportion, hash_key)
    train_loader, val_loader, counts = _finish(train, val, batch_size, sampler, None, placement, device)
    assert val_loader is not None
    return train_loader, val_loader, counts


def load_cifar10_test_data(
    data_dir: Path, batch_size: int, sampler: LabelBasedSampler | None = None, placement: str = "host",
    device: torch.device | str | None = None,
) -> tuple[BatchedTensorLoader, dict[str, int]]:
    log(INFO, f"Data directory: {data_dir!s}")
    data, targets = get_cifar10_data_and_target_tensors(data_dir, False)
    test_set = _make_dataset(data, targets, None, None, cifar10_batch_transform)
    if sampler is not None:
        test_set = sampler.subsample(test_set)
    return BatchedTensorLoader(test_set, batch_size, placement=placement, device=device), {"eval_set": len(test_set)}


def load_msd_dataset(data_path: str, msd_dataset_name: str) -> None:
    """Medical Segmentation Decathlon download helper: needs network access + MONAI; not available offline."""
    try:
        from monai.apps.utils import download_and_extract  # type: ignore[import-not-found]
    except ImportError as exc:
        raise RuntimeError("load_msd_dataset needs the optional 'monai' package and network access") from exc
    from fl4health_b200.utils.msd_dataset_sources import get_msd_dataset_enum, msd_md5_hashes, msd_urls

    enum_value = get_msd_dataset_enum(msd_dataset_name)
    download_and_extract(url=msd_urls[enum_value], output_dir=data_path, hash_val=msd_md5_hashes[enum_value], hash_type="md5")
