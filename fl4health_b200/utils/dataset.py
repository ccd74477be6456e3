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


def _then(first: Callable | None, second: Callable) -> Callable:
    """``second`` applied after ``first`` (or alone when there is no ``first``)."""
    if first is None:
        return second
    return lambda *args: second(first(*args))


class BaseDataset(ABC, Dataset):
    def __init__(self, transform: Callable | None, target_transform: Callable | None) -> None:
        self.transform, self.target_transform = transform, target_transform

    def update_transform(self, f: Callable) -> None:
        """Append ``f`` to the input transform chain."""
        self.transform = _then(self.transform, f)

    def update_target_transform(self, g: Callable) -> None:
        self.target_transform = _then(self.target_transform, g)

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
        self.data, self.targets = data, targets
        self.batch_transform = batch_transform  # applied to a whole [B, ...] batch (vectorised path)

    def __len__(self) -> int:
        return len(self.data)

    def __getitem__(self, index: int) -> tuple[torch.Tensor, torch.Tensor]:
        assert self.targets is not None
        sample, label = self.data[index], self.targets[index]
        sample = sample if self.transform is None else self.transform(sample)
        label = label if self.target_transform is None else self.target_transform(label)
        return sample, label

    def apply_transforms(self, data: torch.Tensor, target: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        """Transforms of an already gathered ``[B, ...]`` batch (the loader gathers whole epochs at once): the batched
        transform if one was given, else the per-sample transform row by row."""
        if self.batch_transform is not None:
            data = self.batch_transform(data)
        elif self.transform is not None:
            data = torch.stack([self.transform(row) for row in data])
        if self.target_transform is not None:
            target = torch.stack([torch.as_tensor(self.target_transform(label)) for label in target])
        return data, target

    def get_batch(self, indices: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        assert self.targets is not None
        return self.apply_transforms(self.data.index_select(0, indices), self.targets.index_select(0, indices))


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
        assert self.target_transform is not None, "Target transform cannot be None."
        view = self.data[index] if self.transform is None else self.transform(self.data[index])
        return view, self.target_transform(view)

    def get_batch(self, indices: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        views, targets = zip(*(self[int(i)] for i in indices))
        return torch.stack(views), torch.stack(targets)


class DictionaryDataset(Dataset):
    def __init__(self, data: dict[str, list[torch.Tensor]], targets: torch.Tensor) -> None:
        self.data, self.targets = data, targets

    def __len__(self) -> int:
        return len(next(iter(self.data.values())))

    def __getitem__(self, index: int) -> tuple[dict[str, torch.Tensor], torch.Tensor]:
        return {field: column[index] for field, column in self.data.items()}, self.targets[index]


class SyntheticDataset(TensorDataset):
    def __init__(self, data: torch.Tensor, targets: torch.Tensor) -> None:
        assert data.shape[0] == targets.shape[0]
        super().__init__(data, targets)


D = TypeVar("D", TensorDataset, DictionaryDataset)


def select_by_indices(dataset: D, selected_indices: torch.Tensor) -> D:
    """A dataset of the same kind restricted to ``selected_indices`` (transforms are shared, tensors are gathered)."""
    if isinstance(dataset, DictionaryDataset):
        rows = [int(i) for i in selected_indices]
        picked = {field: [column[i] for i in rows] for field, column in dataset.data.items()}
        return cast(D, DictionaryDataset(picked, dataset.targets[selected_indices]))
    if not isinstance(dataset, TensorDataset):
        raise TypeError("Dataset type is not supported by this function.")
    subset = copy.copy(dataset)
    subset.data = dataset.data[selected_indices]
    subset.targets = None if dataset.targets is None else dataset.targets[selected_indices]
    return cast(D, subset)
