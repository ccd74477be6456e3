"""Attribute snapshotters used by the state checkpointers (parity: ``fl4health/utils/snapshotter.py:20-281``).

A snapshotter turns ``{key: live object}`` into something ``torch.save``-able and restores it in place.  Scalars are
handled by wrapping them as ``{"None": value}`` in the state checkpointer.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from enum import Enum
from typing import Any, Generic, TypeVar

import torch
from torch import nn
from torch.optim import Optimizer
from torch.optim.lr_scheduler import LRScheduler

T = TypeVar("T")


class AbstractSnapshotter(ABC, Generic[T]):
    @abstractmethod
    def save_attribute(self, attribute: dict[str, T]) -> dict[str, Any]:
        raise NotImplementedError

    @abstractmethod
    def load_attribute(self, attribute_snapshot: dict[str, Any], attribute: dict[str, T]) -> None:
        raise NotImplementedError


def _cpu_tree(obj: Any) -> Any:
    """Detach + move every tensor in a nested container to the CPU (state files must not pin GPU storages)."""
    if isinstance(obj, torch.Tensor):
        return obj.detach().cpu().clone()
    if isinstance(obj, dict):
        return {k: _cpu_tree(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_cpu_tree(v) for v in obj)
    return obj


class OptimizerSnapshotter(AbstractSnapshotter[Optimizer]):
    """Saves only ``state_dict()["state"]`` (moments, step counts); param groups come from the live optimizer."""

    def save_attribute(self, attribute: dict[str, Optimizer]) -> dict[str, Any]:
        return {key: _cpu_tree(opt.state_dict()["state"]) for key, opt in attribute.items()}

    def load_attribute(self, attribute_snapshot: dict[str, Any], attribute: dict[str, Optimizer]) -> None:
        for key, opt in attribute.items():
            full = opt.state_dict()
            full["state"] = attribute_snapshot[key]
            opt.load_state_dict(full)


class LRSchedulerSnapshotter(AbstractSnapshotter[LRScheduler]):
    def save_attribute(self, attribute: dict[str, LRScheduler]) -> dict[str, Any]:
        return {key: sched.state_dict() for key, sched in attribute.items()}

    def load_attribute(self, attribute_snapshot: dict[str, Any], attribute: dict[str, LRScheduler]) -> None:
        for key, sched in attribute.items():
            sched.load_state_dict(attribute_snapshot[key])


class TorchModuleSnapshotter(AbstractSnapshotter[nn.Module]):
    def save_attribute(self, attribute: dict[str, nn.Module]) -> dict[str, Any]:
        return {key: _cpu_tree(dict(model.state_dict())) for key, model in attribute.items()}

    def load_attribute(self, attribute_snapshot: dict[str, Any], attribute: dict[str, nn.Module]) -> None:
        for key, model in attribute.items():
            model.load_state_dict(attribute_snapshot[key])


class _PassThroughSnapshotter(AbstractSnapshotter[T]):
    """Objects stored as-is (pickled) and swapped back wholesale on load."""

    def save_attribute(self, attribute: dict[str, T]) -> dict[str, Any]:
        return dict(attribute)

    def load_attribute(self, attribute_snapshot: dict[str, Any], attribute: dict[str, T]) -> None:
        for key in list(attribute.keys()):
            attribute[key] = attribute_snapshot[key]


class SerializableObjectSnapshotter(_PassThroughSnapshotter[Any]):
    """MetricManager / LossMeter / ReportsManager."""


class SingletonSnapshotter(_PassThroughSnapshotter[Any]):
    """int / float / bool."""


class HistorySnapshotter(_PassThroughSnapshotter[Any]):
    pass


class StringSnapshotter(_PassThroughSnapshotter[str]):
    pass


class BytesSnapshotter(_PassThroughSnapshotter[bytes]):
    pass


class EnumSnapshotter(_PassThroughSnapshotter[Enum]):
    pass
