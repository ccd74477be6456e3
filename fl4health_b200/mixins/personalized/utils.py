"""``ensure_protocol_compliance`` (parity: ``fl4health/mixins/personalized/utils.py:9-31``, there built on ``wrapt``):
decorator for mixin methods that only make sense on a flexible client.  The check is the method part of
``FlexibleClientProtocol`` (``core_protocols.Contract``), done once per class and remembered; the error names what
the class lacks."""

from __future__ import annotations

import functools
from collections.abc import Callable
from typing import Any, TypeVar

from fl4health_b200.mixins.core_protocols import FlexibleClientProtocol

F = TypeVar("F", bound=Callable[..., Any])
_compliant_classes: set[type] = set()


def ensure_protocol_compliance(func: F) -> F:
    @functools.wraps(func)
    def checked(self: Any, *args: Any, **kwargs: Any) -> Any:
        owner = type(self)
        if owner not in _compliant_classes:
            FlexibleClientProtocol.require(owner)  # TypeError("Protocol requirements not met. ...")
            _compliant_classes.add(owner)
        return func(self, *args, **kwargs)

    return checked  # type: ignore[return-value]
