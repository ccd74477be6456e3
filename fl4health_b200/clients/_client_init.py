"""Shared constructor plumbing so algorithm clients do not each restate the nine base arguments."""

from __future__ import annotations

from typing import Any

BASE_KEYS = (
    "data_path", "metrics", "device", "loss_meter_type", "checkpoint_and_state_module", "reporters", "progress_bar",
    "client_name", "engine_options",
)


def split_base_kwargs(kwargs: dict[str, Any]) -> tuple[dict[str, Any], dict[str, Any]]:
    base = {k: v for k, v in kwargs.items() if k in BASE_KEYS}
    extra = {k: v for k, v in kwargs.items() if k not in BASE_KEYS}
    return base, extra
