"""Client-side helpers (parity: ``fl4health/utils/client.py:24-179``)."""

from __future__ import annotations

import copy
from collections.abc import Iterable
from logging import INFO, WARNING
from typing import Any, TypeVar

import torch
from torch import nn

from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config, Scalar
from fl4health_b200.metrics.base_metrics import MetricPrefix
from fl4health_b200.utils.config import narrow_dict_type
from fl4health_b200.utils.logging import LoggingMode
from fl4health_b200.utils.typing import TorchInputType, TorchTargetType

T = TypeVar("T", TorchInputType, TorchTargetType)


def fold_loss_dict_into_metrics(
    metrics: dict[str, Scalar], loss_dict: dict[str, float], logging_mode: LoggingMode
) -> None:
    prefix = MetricPrefix.VAL_PREFIX if logging_mode is LoggingMode.VALIDATION else MetricPrefix.TEST_PREFIX
    metrics.update({f"{prefix.value} {key}": value for key, value in loss_dict.items()})


def set_pack_losses_with_val_metrics(config: Config) -> bool:
    pack = config.get("pack_losses_with_val_metrics", False)
    pack = pack if isinstance(pack, bool) else False  # anything but a real boolean counts as "not requested"
    if pack:
        log(INFO, "As specified in the config, all validation losses will be packed into validation metrics")
    return pack


def move_data_to_device(data: T, device: torch.device) -> T:
    """Async H2D when the source is pinned (the reference's ``.to(device)`` is a blocking copy)."""
    if isinstance(data, torch.Tensor):
        return data if data.device == device else data.to(device, non_blocking=True)
    if isinstance(data, dict):
        return {k: (v if v.device == device else v.to(device, non_blocking=True)) for k, v in data.items()}
    raise TypeError("data must be of type torch.Tensor or dict[str, torch.Tensor].")


def check_if_batch_is_empty_and_verify_input(input: TorchInputType) -> bool:
    if isinstance(input, torch.Tensor):
        return len(input) == 0
    if isinstance(input, dict):
        lengths = {len(v) for v in input.values()}
        if len(lengths) != 1:
            raise ValueError("Not all tensors in the dictionary have the same size.")
        return lengths.pop() == 0
    raise TypeError("Input must be of type torch.Tensor or dict[str, torch.Tensor].")


def clone_and_freeze_model(model: nn.Module) -> nn.Module:
    cloned = copy.deepcopy(model)
    for param in cloned.parameters():
        param.requires_grad = False
        param.grad = None
    cloned.eval()
    return cloned


def maybe_progress_bar(iterable: Iterable, display_progress_bar: bool) -> Iterable:
    if not display_progress_bar:
        return iterable
    try:
        from tqdm import tqdm
    except ImportError:  # pragma: no cover
        return iterable
    kwargs: Any = {"leave": True, "ascii": " >=", "unit": "steps", "dynamic_ncols": True}
    return tqdm(iterable, **kwargs)


def process_and_check_validation_steps(config: Config, val_loader: Any) -> int | None:
    if "num_validation_steps" not in config:
        return None
    log(
        INFO,
        "num_validation_steps specified in config. Only a subset of batches will be processed from the validation "
        "set during evaluation.",
    )
    num_validation_steps = narrow_dict_type(config, "num_validation_steps", int)
    assert num_validation_steps > 0, "num_validation_steps must not be 0"
    loader_len = len(val_loader)
    assert loader_len > 0, "Dataloader must have length greater than 0."
    if num_validation_steps > loader_len:
        log(WARNING, f"num_validation_steps: {num_validation_steps} is larger than the validation dataloader: {loader_len}")
    return num_validation_steps
