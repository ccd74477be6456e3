"""``ensure_protocol_compliance`` (parity: ``fl4health/mixins/personalized/utils.py:9-31``, there built on ``wrapt``):
decorator for mixin methods that only make sense on a ``FlexibleClient``."""

from __future__ import annotations

import functools
from collections.abc import Callable
from typing import Any, TypeVar

F = TypeVar("F", bound=Callable[..., Any])


def ensure_protocol_compliance(func: F) -> F:
    @functools.wraps(func)
    def wrapper(self: Any, *args: Any, **kwargs: Any) -> Any:
        from fl4health_b200.clients.flexible.base import FlexibleClient

        if not isinstance(self, FlexibleClient):
            raise TypeError("Protocol requirements not met.")
        return func(self, *args, **kwargs)

    return wrapper  # type: ignore[return-value]
