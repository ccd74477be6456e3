"""Dataset containers (parity: ``fl4health/utils/dataset.py:12-243``).

``TensorDataset`` keeps the per-sample ``__getitem__`` contract for stock ``DataLoader`` use, and additionally offers
``get_batch(indices)`` — a vectorised gather that applies *batched* transforms — which the engine's
``BatchedTensorLoader`` uses to avoid per-sample Python work (per-sample transform in ``__getitem__``).
"""

from __future__ import annotations

import copy
from abc import ABC, abstractmethod
from collections.abc import Callable
from typing import TypeVar, cast

import torch
from torch.utils.data import Dataset


class BaseDataset(ABC, Dataset):
    def __init__(self, transform: Callable | None, target_transform: Callable | None) -> None:
        self.transform = transform
        self.target_transform = target_transform

    def update_transform(self, f: Callable) -> None:
        previous = self.transform
        self.transform = (lambda *x: f(previous(*x))) if previous else f

    def update_target_transform(self, g: Callable) -> None:
        previous = self.target_transform
        self.target_transform = (lambda *x: g(previous(*x))) if previous else g

    @abstractmethod
    def __getitem__(self, index: int) -> tuple[torch.Tensor, torch.Tensor]:
        raise NotImplementedError

    @abstractmethod
    def __len__(self) -> int:
        raise NotImplementedError


class TensorDataset(BaseDataset):
    def __init__(
        self,
        data: torch.Tensor,
        targets: torch.Tensor | None = None,
        transform: Callable | None = None,
        target_transform: Callable | None = None,
        batch_transform: Callable | None = None,
    ) -> None:
        super().__init__(transform, target_transform)
        self.data = data
        self.targets = targets
        self.batch_transform = batch_transform  # applied to a whole [B, ...] batch (vectorised path)

    def __getitem__(self, index: int) -> tuple[torch.Tensor, torch.Tensor]:
        assert self.targets is not None
        data, target = self.data[index], self.targets[index]
        if self.transform is not None:
            data = self.transform(data)
        if self.target_transform is not None:
            target = self.target_transform(target)
        return data, target

    def get_batch(self, indices: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        assert self.targets is not None
        return self.apply_transforms(self.data.index_select(0, indices), self.targets.index_select(0, indices))

    def apply_transforms(self, data: torch.Tensor, target: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        """Transforms of an already gathered ``[B, ...]`` batch (the loader gathers whole epochs at once)."""
        if self.batch_transform is not None:
            data = self.batch_transform(data)
        elif self.transform is not None:
            data = torch.stack([self.transform(sample) for sample in data])
        if self.target_transform is not None:
            target = torch.stack([torch.as_tensor(self.target_transform(t)) for t in target])
        return data, target

    def __len__(self) -> int:
        return len(self.data)


class SslTensorDataset(TensorDataset):
    """Self-supervised pairs: the target is a transformed view of the input."""

    def __init__(
        self,
        data: torch.Tensor,
        targets: torch.Tensor | None = None,
        transform: Callable | None = None,
        target_transform: Callable | None = None,
    ) -> None:
        assert targets is None, "SslTensorDataset targets must be None"
        super().__init__(data, targets, transform, target_transform)

    def __getitem__(self, index: int) -> tuple[torch.Tensor, torch.Tensor]:
        data = self.data[index]
        assert self.target_transform is not None, "Target transform cannot be None."
        if self.transform is not None:
            data = self.transform(data)
        return data, self.target_transform(data)

    def get_batch(self, indices: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        pairs = [self[int(i)] for i in indices]
        return torch.stack([p[0] for p in pairs]), torch.stack([p[1] for p in pairs])


class DictionaryDataset(Dataset):
    def __init__(self, data: dict[str, list[torch.Tensor]], targets: torch.Tensor) -> None:
        self.data = data
        self.targets = targets

    def __getitem__(self, index: int) -> tuple[dict[str, torch.Tensor], torch.Tensor]:
        return {key: val[index] for key, val in self.data.items()}, self.targets[index]

    def __len__(self) -> int:
        return len(next(iter(self.data.values())))


class SyntheticDataset(TensorDataset):
    def __init__(self, data: torch.Tensor, targets: torch.Tensor) -> None:
        assert data.shape[0] == targets.shape[0]
        super().__init__(data, targets)


D = TypeVar("D", TensorDataset, DictionaryDataset)


def select_by_indices(dataset: D, selected_indices: torch.Tensor) -> D:
    if isinstance(dataset, TensorDataset):
        subset = copy.copy(dataset)
        subset.data = dataset.data[selected_indices]
        if dataset.targets is not None:
            subset.targets = dataset.targets[selected_indices]
        return cast(D, subset)
    if isinstance(dataset, DictionaryDataset):
        new_data = {key: [val[int(i)] for i in selected_indices] for key, val in dataset.data.items()}
        return cast(D, DictionaryDataset(new_data, dataset.targets[selected_indices]))
    raise TypeError("Dataset type is not supported by this function.")
