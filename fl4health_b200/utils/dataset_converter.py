"""Dataset converters (parity: ``fl4health/utils/dataset_converter.py:9-226``): wrap a ``TensorDataset`` so that
``(data, target)`` is re-shaped for auto-encoder training — target := data, and for conditional AEs the condition
(label or fixed vector) is packed onto the flattened input.  Both the per-sample (``__getitem__``) and the vectorised
(``get_batch``) access paths are converted."""

from __future__ import annotations

from collections.abc import Callable
from functools import partial

import torch

from fl4health_b200.utils.dataset import TensorDataset


class DatasetConverter(TensorDataset):
    def __init__(
        self, converter_function: Callable[[torch.Tensor, torch.Tensor], tuple[torch.Tensor, torch.Tensor]],
        dataset: TensorDataset | None,
    ) -> None:
        assert dataset is None or dataset.targets is not None
        self.converter_function = converter_function
        self.dataset = dataset
        self.transform = None
        self.target_transform = None
        self.batch_transform = None

    # the wrapped dataset's tensors stay reachable (samplers / loaders look at .data / .targets)
    @property
    def data(self) -> torch.Tensor:  # type: ignore[override]
        assert self.dataset is not None
        return self.dataset.data

    @data.setter
    def data(self, value: torch.Tensor) -> None:
        assert self.dataset is not None
        self.dataset.data = value

    @property
    def targets(self) -> torch.Tensor | None:  # type: ignore[override]
        assert self.dataset is not None
        return self.dataset.targets

    @targets.setter
    def targets(self, value: torch.Tensor | None) -> None:
        assert self.dataset is not None
        self.dataset.targets = value

    def __getitem__(self, index: int) -> tuple[torch.Tensor, torch.Tensor]:
        assert self.dataset is not None, "Error: no dataset is set, use convert_dataset(your_dataset: TensorDataset)"
        data, target = self.dataset[index]
        return self.converter_function(data, target)

    def get_batch(self, indices: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        assert self.dataset is not None
        data, target = self.dataset.get_batch(indices)
        pairs = [self.converter_function(d, t) for d, t in zip(data, target)]
        return torch.stack([p[0] for p in pairs]), torch.stack([p[1] for p in pairs])

    def __len__(self) -> int:
        assert self.dataset is not None, "Error: dataset is should be either converted or initiated."
        return len(self.dataset)

    def convert_dataset(self, dataset: TensorDataset) -> TensorDataset:
        self.dataset = dataset
        return self


class AutoEncoderDatasetConverter(DatasetConverter):
    def __init__(
        self, condition: str | torch.Tensor | None = None, do_one_hot_encoding: bool = False,
        custom_converter_function: Callable | None = None, condition_vector_size: int | None = None,
    ) -> None:
        """``condition``: None (plain AE / VAE), ``"label"`` (condition each sample on its target) or a fixed 1-D
        tensor (same condition for every sample of this client)."""
        self.condition = condition
        if isinstance(condition, torch.Tensor):
            assert condition.dim() == 1, f"Error: condition should be a 1D vector instead of {condition.dim()}D tensor."
        self.data_shape: torch.Size
        self.do_one_hot_encoding = do_one_hot_encoding
        self.condition_vector_size = condition_vector_size
        if custom_converter_function is None:
            function = self._setup_converter_function()
        else:
            assert condition_vector_size is not None, "Error: The condition should be specified for a custom converter function."
            function = custom_converter_function
        super().__init__(function, dataset=None)

    def convert_dataset(self, dataset: TensorDataset) -> TensorDataset:
        assert dataset.targets is not None
        self.dataset = dataset
        self.data_shape = dataset[0][0].shape  # shape after the dataset's own transforms: needed to unpack
        return self

    def get_condition_vector_size(self) -> int:
        if isinstance(self.condition, str) and self.condition == "label":
            assert self.dataset is not None and self.dataset.targets is not None, "Error: no dataset is passed to the converter."
            if self.do_one_hot_encoding:
                return len(torch.unique(self.dataset.targets))
            return len(self.dataset.targets[0])
        if isinstance(self.condition, torch.Tensor):
            return self.condition.size(0)
        if self.condition_vector_size is not None:
            return self.condition_vector_size
        if self.condition is None:
            return 0
        raise NotImplementedError("Error: support for this type of condition is not added to the data converter.")

    def _setup_converter_function(self) -> Callable:
        if self.condition is None:
            return self._only_replace_target_with_data
        if isinstance(self.condition, str) and self.condition == "label":
            return self._cat_input_label
        if isinstance(self.condition, torch.Tensor):
            return self._cat_input_condition
        raise NotImplementedError("Error: support for this type of condition is not added.")

    def _only_replace_target_with_data(self, data: torch.Tensor, target: torch.Tensor | None) -> tuple[torch.Tensor, torch.Tensor]:  # noqa: ARG002
        return data, data

    def _cat_input_condition(self, data: torch.Tensor, target: torch.Tensor | None) -> tuple[torch.Tensor, torch.Tensor]:  # noqa: ARG002
        assert isinstance(self.condition, torch.Tensor), "Error: condition should be a torch tensor"
        return torch.cat([data.reshape(-1), self.condition.to(data.dtype)]), data

    def _cat_input_label(self, data: torch.Tensor, target: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        if self.do_one_hot_encoding:
            target = torch.nn.functional.one_hot(target, num_classes=self.get_condition_vector_size())
        return torch.cat([data.reshape(-1), target.reshape(-1).to(data.dtype)]), data

    def get_unpacking_function(self) -> Callable[[torch.Tensor], tuple[torch.Tensor, torch.Tensor]]:
        return partial(
            AutoEncoderDatasetConverter.unpack_input_condition, cond_vec_size=self.get_condition_vector_size(),
            data_shape=self.data_shape,
        )

    @staticmethod
    def unpack_input_condition(packed_data: torch.Tensor, cond_vec_size: int, data_shape: torch.Size) -> tuple[torch.Tensor, torch.Tensor]:
        """``[B, prod(data_shape) + cond]`` -> (``[B, *data_shape]``, ``[B, cond]``)."""
        assert data_shape is not None
        split = packed_data.shape[1] - cond_vec_size
        return packed_data[:, :split].reshape(-1, *data_shape), packed_data[:, split:]
