"""Layer / parameter selection rules for partial exchange.

Parity: ``fl4health/parameter_exchange/parameter_selection_criteria.py:13-267``.  All scoring stays on the device;
drift norms of all layers are gathered into one tensor and read back with a single sync (the reference calls
``.item()`` per layer).  FedPM masks are sampled with ``torch.bernoulli`` on the device instead of
``scipy.stats.bernoulli.rvs`` on the CPU (:202-205).
"""

from __future__ import annotations

import math
from functools import partial

import torch
from torch import Tensor, nn

from fl4health_b200.common.typing import NDArrays
from fl4health_b200.utils.typing import LayerSelectionFunction


class LayerSelectionFunctionConstructor:
    def __init__(
        self, norm_threshold: float, exchange_percentage: float, normalize: bool = True, select_drift_more: bool = True
    ) -> None:
        assert 0 < exchange_percentage <= 1
        assert norm_threshold > 0
        self.norm_threshold = norm_threshold
        self.exchange_percentage = exchange_percentage
        self.normalize = normalize
        self.select_drift_more = select_drift_more

    def select_by_threshold(self) -> LayerSelectionFunction:
        return partial(select_layers_by_threshold, self.norm_threshold, self.normalize, self.select_drift_more)

    def select_by_percentage(self) -> LayerSelectionFunction:
        return partial(select_layers_by_percentage, self.exchange_percentage, self.normalize, self.select_drift_more)


def _drift_norms(model: nn.Module, initial_model: nn.Module, normalize: bool) -> tuple[dict[str, Tensor], list[float]]:
    states, initial = model.state_dict(), initial_model.state_dict()
    norms = []
    for name, value in states.items():
        diff = (value - initial[name]).float()
        norm = torch.linalg.norm(diff)
        norms.append(norm / max(diff.numel(), 1) if normalize else norm)
    values = torch.stack(norms).tolist() if norms else []  # one device->host transfer for all layers
    return states, values


def _calculate_drift_norm(t1: Tensor, t2: Tensor, normalize: bool) -> float:
    diff = (t1 - t2).float()
    norm = torch.linalg.norm(diff)
    return float((norm / diff.numel() if normalize else norm).item())


def select_layers_by_threshold(
    threshold: float, normalize: bool, select_drift_more: bool, model: nn.Module, initial_model: nn.Module
) -> tuple[NDArrays, list[str]]:
    states, norms = _drift_norms(model, initial_model, normalize)
    names = [
        name
        for name, norm in zip(states.keys(), norms)
        if (norm > threshold if select_drift_more else norm <= threshold)
    ]
    return NDArrays([states[n].detach() for n in names]), names


def select_layers_by_percentage(
    exchange_percentage: float, normalize: bool, select_drift_more: bool, model: nn.Module, initial_model: nn.Module
) -> tuple[NDArrays, list[str]]:
    states, norms = _drift_norms(model, initial_model, normalize)
    by_name = dict(zip(states.keys(), norms))
    count = int(math.ceil(len(by_name) * exchange_percentage))
    names = sorted(by_name.keys(), key=lambda n: by_name[n], reverse=select_drift_more)[:count]
    return NDArrays([states[n].detach() for n in names]), names


# ---- score generators for sparse (COO) exchange ---------------------------------------------------------------
def _paired(model: nn.Module, initial_model: nn.Module | None):  # noqa: ANN202
    assert initial_model is not None
    initial = initial_model.state_dict()
    for name, current in model.state_dict().items():
        yield name, current, initial[name]


def largest_final_magnitude_scores(model: nn.Module, initial_model: nn.Module | None) -> dict[str, Tensor]:
    return {name: value.abs() for name, value in model.state_dict().items()}


def smallest_final_magnitude_scores(model: nn.Module, initial_model: nn.Module | None) -> dict[str, Tensor]:
    return {name: -value.abs() for name, value in model.state_dict().items()}


def largest_magnitude_change_scores(model: nn.Module, initial_model: nn.Module | None) -> dict[str, Tensor]:
    return {name: (cur - old).abs() for name, cur, old in _paired(model, initial_model)}


def smallest_magnitude_change_scores(model: nn.Module, initial_model: nn.Module | None) -> dict[str, Tensor]:
    return {name: -(cur - old).abs() for name, cur, old in _paired(model, initial_model)}


def largest_increase_in_magnitude_scores(model: nn.Module, initial_model: nn.Module | None) -> dict[str, Tensor]:
    return {name: cur.abs() - old.abs() for name, cur, old in _paired(model, initial_model)}


def smallest_increase_in_magnitude_scores(model: nn.Module, initial_model: nn.Module | None) -> dict[str, Tensor]:
    return {name: -(cur.abs() - old.abs()) for name, cur, old in _paired(model, initial_model)}


# ---- FedPM ----------------------------------------------------------------------------------------------------
def _sample_masks(score_tensor: Tensor) -> Tensor:
    """Binary mask ~ Bernoulli(sigmoid(score)), uint8 on the score's device."""
    return torch.bernoulli(torch.sigmoid(score_tensor.float())).to(torch.uint8)


def _process_masked_module(
    module: nn.Module, model_state_dict: dict[str, Tensor], module_name: str | None = None
) -> tuple[NDArrays, list[str]]:
    prefix = f"{module_name}." if module_name else ""
    names = [f"{prefix}weight_scores"]
    if "bias_scores" in module.state_dict():
        names.append(f"{prefix}bias_scores")
    return NDArrays([_sample_masks(model_state_dict[n]) for n in names]), names


def select_scores_and_sample_masks(model: nn.Module, initial_model: nn.Module | None) -> tuple[NDArrays, list[str]]:
    from fl4health_b200.model_bases.masked_layers.masked_layers_utils import is_masked_module

    states = model.state_dict()
    with torch.no_grad():
        if is_masked_module(model):
            return _process_masked_module(model, states)
        masks, names = NDArrays(), []
        for name, module in model.named_modules():
            if is_masked_module(module):
                module_masks, module_names = _process_masked_module(module, states, name)
                masks.extend(module_masks)
                names.extend(module_names)
        return masks, names
