"""Glue between models and the in-house DP engine (parity: ``fl4health/utils/privacy_utilities.py:11-91``)."""

from __future__ import annotations

from logging import INFO, WARNING
from typing import Any

from torch import nn

from fl4health_b200.common.logger import log
from fl4health_b200.privacy.dp_engine import GradSampleModule, ModuleValidator, wrap_model


def privacy_validate_and_fix_modules(model: nn.Module) -> tuple[nn.Module, bool]:
    """Replace DP-incompatible layers (BatchNorm -> GroupNorm).  Returns the model and whether parameters changed
    (in which case optimizers must be rebuilt)."""
    errors = ModuleValidator.validate(model, strict=False)
    fixable = [e for e in errors if "BatchNorm" in e]
    reinitialize_optimizer = len(fixable) > 0
    if fixable:
        log(WARNING, "Found layers that do not comply with DP training; they will be replaced with DP compliant layers.")
        for error in fixable:
            log(WARNING, f"DP validation error: {error}")
        model = ModuleValidator.fix(model)
    remaining = ModuleValidator.validate(model, strict=False)
    if remaining:
        raise ValueError("Model cannot be made DP-compatible:\n" + "\n".join(remaining))
    return model, reinitialize_optimizer


def convert_model_to_opacus_model(model: nn.Module, grad_sample_mode: str = "hooks", *args: Any, **kwargs: Any) -> GradSampleModule:
    if isinstance(model, GradSampleModule):
        log(INFO, f"Provided model is already of type {type(model)}, skipping conversion")
        return model
    return wrap_model(model, grad_sample_mode, *args, **kwargs)


def map_model_to_opacus_model(model: nn.Module, grad_sample_mode: str = "hooks", *args: Any, **kwargs: Any) -> GradSampleModule:
    model, _ = privacy_validate_and_fix_modules(model)
    return convert_model_to_opacus_model(model, grad_sample_mode, *args, **kwargs)
