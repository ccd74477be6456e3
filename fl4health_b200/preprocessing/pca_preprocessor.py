"""Dimensionality reduction with a federated PCA checkpoint (parity: ``fl4health/preprocessing/pca_preprocessor.py``)."""

from __future__ import annotations

from functools import partial
from pathlib import Path

import torch

from fl4health_b200.model_bases.pca import PcaModule
from fl4health_b200.utils.dataset import TensorDataset


class PcaPreprocessor:
    def __init__(self, checkpointing_path: Path) -> None:
        self.checkpointing_path = checkpointing_path
        self.pca_module: PcaModule = self.load_pca_module()

    def load_pca_module(self) -> PcaModule:
        module = torch.load(self.checkpointing_path, weights_only=False)
        module.eval()
        return module

    def reduce_dimension(self, new_dimension: int, dataset: TensorDataset) -> TensorDataset:
        """Appends the projection onto the top ``new_dimension`` components to the dataset's transforms (the projection
        is a matmul: it is installed as a *batched* transform when the dataset has no per-sample transform)."""
        projection = partial(self.pca_module.project_lower_dim, k=new_dimension)
        if dataset.transform is None and getattr(dataset, "batch_transform", None) is None:
            dataset.batch_transform = projection
            dataset.transform = lambda sample: projection(sample.unsqueeze(0)).squeeze(0)
        else:
            dataset.update_transform(projection)
        return dataset
