"""Label-based subsamplers that induce heterogeneity (parity: ``fl4health/utils/sampler.py:18-183``)."""

from __future__ import annotations

import math
from abc import ABC, abstractmethod
from collections.abc import Set
from logging import INFO, WARNING
from typing import Any, TypeVar

import numpy as np
import torch

from fl4health_b200.common.logger import log
from fl4health_b200.utils.dataset import DictionaryDataset, TensorDataset, select_by_indices

T = TypeVar("T")
D = TypeVar("D", TensorDataset, DictionaryDataset)


class LabelBasedSampler(ABC):
    def __init__(self, unique_labels: list[Any]) -> None:
        self.unique_labels = unique_labels
        self.num_classes = len(unique_labels)

    @abstractmethod
    def subsample(self, dataset: D) -> D:
        raise NotImplementedError


class MinorityLabelBasedSampler(LabelBasedSampler):
    """Keeps every sample of the majority labels and a ``downsampling_ratio`` fraction of each minority label."""

    def __init__(self, unique_labels: list[T], downsampling_ratio: float, minority_labels: Set[T]) -> None:
        super().__init__(unique_labels)
        self.downsampling_ratio = downsampling_ratio
        self.minority_labels = minority_labels

    def subsample(self, dataset: D) -> D:
        assert dataset.targets is not None, "A label-based sampler requires targets but this dataset has no targets"
        chosen = []
        for label in self.unique_labels:
            indices = (dataset.targets == label).nonzero().reshape(-1)
            if label in self.minority_labels:
                indices = self._get_random_subsample(indices, int(indices.shape[0] * self.downsampling_ratio))
            chosen.append(indices)
        return select_by_indices(dataset, torch.cat(chosen, dim=0))

    def _get_random_subsample(self, tensor_to_subsample: torch.Tensor, subsample_size: int) -> torch.Tensor:
        size = tensor_to_subsample.shape[0]
        assert subsample_size < size
        return tensor_to_subsample[torch.randperm(size)[:subsample_size]]


class DirichletLabelBasedSampler(LabelBasedSampler):
    """Draws class proportions from ``Dirichlet(beta)`` once, then samples (with replacement)
    ``sample_percentage * len(dataset)`` points following them.  Small beta -> more heterogeneous clients."""

    def __init__(self, unique_labels: list[Any], hash_key: int | None = None, sample_percentage: float = 0.5, beta: float = 100) -> None:
        super().__init__(unique_labels)
        self.hash_key = hash_key
        self.torch_generator: torch.Generator | None = None
        if hash_key is not None:
            log(INFO, f"Setting seed to {hash_key} for the Torch and Numpy Generators")
            log(WARNING, "Note that setting a hash key here will override any torch and numpy seeds that you have set")
            self.torch_generator = torch.Generator().manual_seed(hash_key)
            self.probabilities = np.random.default_rng(hash_key).dirichlet(np.repeat(beta, self.num_classes))
        else:
            self.probabilities = np.random.dirichlet(np.repeat(beta, self.num_classes))
        log(INFO, f"Setting probabilities to {self.probabilities}")
        self.sample_percentage = sample_percentage

    def subsample(self, dataset: D) -> D:
        assert dataset.targets is not None, "A label-based sampler requires targets but this dataset has no targets"
        assert self.sample_percentage <= 1.0
        total = int(len(dataset) * self.sample_percentage)
        picks = []
        for label, prob in zip(self.unique_labels, self.probabilities):
            class_idx = torch.where(dataset.targets == label)[0]
            count = math.ceil(prob * total)
            draws = torch.multinomial(torch.ones(class_idx.size(0)), count, replacement=True, generator=self.torch_generator)
            picks.append(class_idx[draws])
        return select_by_indices(dataset, torch.cat(picks, dim=0).long()[:total])
