"""Mixin root (parity: ``fl4health/mixins/base.py:11-39``): validates at class-creation time that a flexible mixin
is combined with a ``FlexibleClient``."""

from __future__ import annotations

from typing import Any


class BaseFlexibleMixin:
    _is_flexible_mixin = True

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)

    def __init_subclass__(cls, **kwargs: Any) -> None:
        super().__init_subclass__(**kwargs)
        if cls.__dict__.get("_dynamically_created", False):
            return
        from fl4health_b200.clients.flexible.base import FlexibleClient

        names = {base.__name__ for base in cls.__mro__}
        is_pure_mixin = all(getattr(base, "_is_flexible_mixin", False) or base is object for base in cls.__mro__[1:])
        if not is_pure_mixin and not issubclass(cls, FlexibleClient):
            raise RuntimeError(f"Class {cls.__name__} inherits from a flexible mixin but is not a FlexibleClient ({sorted(names)}).")
