"""Dirichlet label-based partitioning of one dataset into client shards (parity:
``fl4health/utils/partitioners.py:16-229``).  For each label the sample indices are split across partitions according
to ``Dirichlet(beta)`` (or a user prior); partitions are disjoint."""

from __future__ import annotations

import math
from logging import INFO, WARNING
from typing import Generic, TypeVar

import numpy as np
import torch

from fl4health_b200.common.logger import log
from fl4health_b200.utils.dataset import DictionaryDataset, TensorDataset, select_by_indices

T = TypeVar("T")
D = TypeVar("D", TensorDataset, DictionaryDataset)


class DirichletLabelBasedAllocation(Generic[T]):
    def __init__(
        self, number_of_partitions: int, unique_labels: list[T], min_label_examples: int | None = None,
        beta: float | None = None, prior_distribution: dict[T, np.ndarray] | None = None,
    ) -> None:
        assert (beta is not None) ^ (prior_distribution is not None), "Either beta or a prior distribution must be provided, but not both."
        self.number_of_partitions = number_of_partitions
        self.unique_labels = unique_labels
        self.n_unique_labels = len(unique_labels)
        self.beta = beta
        self.min_label_examples = min_label_examples or 0
        self.prior_distribution = prior_distribution
        if prior_distribution is not None:
            assert len(prior_distribution) == self.n_unique_labels, "The length of the prior must match the number of labels"
            if self.min_label_examples > 0:
                log(WARNING, "A prior distribution has been provided for the partitioner so min_label_examples will be ignored.")

    def partition_label_indices(self, label: T, label_indices: torch.Tensor) -> tuple[list[torch.Tensor], int, np.ndarray]:
        """Split the indices of one label; returns (per-partition indices, smallest partition size, allocation probs)."""
        if self.prior_distribution is not None:
            prior = np.asarray(self.prior_distribution[label], dtype=np.float64)
            assert len(prior) == self.number_of_partitions, (
                f"The length of the prior distribution for label ({label!s}) must match the number of partitions")
            if prior.sum() != 1:
                log(WARNING, f"The provided prior distribution for label ({label!s}) does not sum to 1. It will be normalized to sum to 1.")
            allocations = prior / prior.sum()
        elif self.beta is not None:
            allocations = np.random.dirichlet(np.repeat(self.beta, self.number_of_partitions))
        else:
            raise ValueError("Either beta or a prior distribution must be provided.")
        log(INFO, f"The allocation distribution for label ({label!s}) is {allocations}")
        total = len(label_indices)
        counts = [math.floor(p * total) for p in allocations]
        smallest = min(counts)
        counts.append(total - sum(counts))  # rounding remainder: an extra, discarded, partition
        shuffled = label_indices[torch.randperm(total)]
        return list(torch.split(shuffled, counts))[:-1], smallest, allocations

    def partition_dataset(self, original_dataset: D, max_retries: int | None = 5) -> tuple[list[D], dict[T, np.ndarray]]:
        targets = original_dataset.targets
        assert targets is not None, "A label-based partitioner requires targets but this dataset has no targets"
        shards = [torch.empty(0, dtype=torch.int64) for _ in range(self.number_of_partitions)]
        probabilities: dict[T, np.ndarray] = {}
        attempts = 0
        for label in self.unique_labels:
            label_indices = torch.where(targets == label)[0]
            while True:
                parts, smallest, allocation = self.partition_label_indices(label, label_indices)
                if self.prior_distribution is not None or smallest >= self.min_label_examples:
                    probabilities[label] = allocation
                    shards = [torch.cat((shard, part)) for shard, part in zip(shards, parts)]
                    break
                attempts += 1
                log(INFO, f"Too few datapoints in a partition. One partition had {smallest} but the minimum requested was "
                          f"{self.min_label_examples}. Resampling the partition...")
                if max_retries is not None and attempts >= max_retries:
                    raise ValueError(f"Max Retries: {max_retries} reached. Partitioning failed to satisfy the minimum label threshold")
        return [select_by_indices(original_dataset, shard) for shard in shards], probabilities
