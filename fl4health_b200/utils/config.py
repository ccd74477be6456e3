"""YAML config loading + typed dictionary access (parity: ``fl4health/utils/config.py:7-116``)."""

from __future__ import annotations

from collections.abc import Callable
from typing import Any, TypeVar

import yaml

REQUIRED_CONFIG: dict[str, type] = {"n_server_rounds": int, "batch_size": int}
T = TypeVar("T")


class InvalidConfigError(ValueError):
    pass


def check_config(config: dict[str, Any]) -> None:
    for key, expected in REQUIRED_CONFIG.items():
        if key not in config:
            raise InvalidConfigError(f"{key} must be specified in Config File")
        if not isinstance(config[key], expected) or isinstance(config[key], bool):
            raise InvalidConfigError(f"{key} must be of type {expected}")
        if config[key] <= 0:
            raise InvalidConfigError(f"{key} must be greater than 0")


def load_config(config_path: str) -> dict[str, Any]:
    with open(config_path) as handle:
        config = yaml.safe_load(handle)
    check_config(config)
    return config


def narrow_dict_type(dictionary: dict[str, Any], key: str, narrow_type_to: type[T]) -> T:
    if key not in dictionary:
        raise ValueError(f"{key} is not present in the Dictionary.")
    value = dictionary[key]
    if not isinstance(value, narrow_type_to):
        raise ValueError(f"Provided key ({key}) value does not have correct type")
    return value


def narrow_dict_type_and_set_attribute(
    self: object,
    dictionary: dict,
    dictionary_key: str,
    attribute_name: str,
    narrow_type_to: type[T],
    func: Callable[[Any], Any] | None = None,
) -> None:
    value: Any = narrow_dict_type(dictionary, dictionary_key, narrow_type_to)
    setattr(self, attribute_name, func(value) if func is not None else value)


def make_dict_with_epochs_or_steps(local_epochs: int | None = None, local_steps: int | None = None) -> dict[str, int]:
    if local_epochs is not None:
        return {"local_epochs": local_epochs}
    if local_steps is not None:
        return {"local_steps": local_steps}
    return {}
