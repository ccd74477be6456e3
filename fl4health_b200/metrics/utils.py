"""Shape alignment helpers for count-based metrics (parity: ``fl4health/metrics/utils.py:4-186``)."""

from __future__ import annotations

import torch


def infer_label_dim(tensor1: torch.Tensor, tensor2: torch.Tensor) -> int:
    """Index (in ``tensor1``) of the label axis, given that ``tensor2`` is the same data with that axis missing or of
    size 1.  Ambiguous cases raise ``AssertionError``."""
    assert tensor1.shape != tensor2.shape, f"Could not infer the label dimension of tensors with the same shape: {tensor1.shape}"
    extra = tensor1.ndim - tensor2.ndim
    assert extra in (0, 1), (
        f"Could not infer the label dimension of tensors with shapes: tensor1: {tensor1.shape}), tensor 2: "
        f"({tensor2.shape}). Expected tensor1 to be larger than tensor2 by at most 1 dimension."
    )
    candidates: list[int] = []
    j = 0
    for i, size in enumerate(tensor1.shape):
        if j < tensor2.ndim and size == tensor2.shape[j]:
            j += 1
            continue
        candidates.append(i)
        if extra == 0:
            j += 1  # same rank: a mismatching axis consumes the matching position of tensor2
    assert len(candidates) == 1, (
        f"Could not infer the label dimension of tensors with shapes: ({tensor1.shape}), ({tensor2.shape}). "
        "Found multiple axes that could be the label dimension."
    )
    dim = candidates[0]
    if extra == 1 and dim > 0:
        assert tensor1.shape[dim] != tensor1.shape[dim - 1], (
            f"Could not infer the label dimension of tensors with shapes: ({tensor1.shape}), ({tensor2.shape}). "
            "A dimension adjacent to the label dimension appears to have the same size."
        )
    if extra == 0:
        assert 1 in (tensor1.shape[dim], tensor2.shape[dim]), (
            f"Could not infer the label dimension of tensors with shapes: ({tensor1.shape}), ({tensor2.shape}). "
            "The inferred candidate dimension has different sizes on each tensor, was expecting one to be empty."
        )
    return dim


def map_label_index_tensor_to_one_hot(label_index_tensor: torch.Tensor, target_shape: torch.Size, label_dim: int) -> torch.Tensor:
    assert label_dim < label_index_tensor.ndim, f"Label dim: {label_dim} too large for target shape: {label_index_tensor.shape}"
    assert label_index_tensor.shape[label_dim] == 1, (
        f"Expected label_dim {label_dim} of label_index_tensor to be of size 1, but got {label_index_tensor.shape[label_dim]}"
    )
    one_hot = torch.zeros(target_shape, device=label_index_tensor.device)
    return one_hot.scatter_(label_dim, label_index_tensor.to(torch.int64), 1)


def align_pred_and_target_shapes(
    preds: torch.Tensor, targets: torch.Tensor, label_dim: int | None = None
) -> tuple[torch.Tensor, torch.Tensor]:
    """If shapes differ, one side is label-index encoded: one-hot it along the (given or inferred) label axis."""
    if preds.shape == targets.shape:
        return preds, targets
    assert abs(preds.ndim - targets.ndim) <= 1, f"Can not align pred and target tensors with shapes {preds.shape}, {targets.shape}"
    if preds.ndim > targets.ndim:
        dim = infer_label_dim(preds, targets) if label_dim is None else label_dim
        return preds, map_label_index_tensor_to_one_hot(targets.unsqueeze(dim), preds.shape, dim)
    if preds.ndim < targets.ndim:
        dim = infer_label_dim(targets, preds) if label_dim is None else label_dim
        return map_label_index_tensor_to_one_hot(preds.unsqueeze(dim), targets.shape, dim), targets
    dim = infer_label_dim(preds, targets) if label_dim is None else label_dim
    if preds.shape[dim] < targets.shape[dim]:
        return map_label_index_tensor_to_one_hot(preds, targets.shape, dim), targets
    return preds, map_label_index_tensor_to_one_hot(targets, preds.shape, dim)
