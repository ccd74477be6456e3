"""Reporter interface (parity: ``fl4health/reporting/base_reporter.py:10-53``)."""

from __future__ import annotations

from typing import Any


class BaseReporter:
    def initialize(self, **kwargs: Any) -> None:
        """Receives identifying information (``id``, ``name``) from the owning client/server."""

    def report(
        self, data: dict, round: int | None = None, epoch: int | None = None, step: int | None = None
    ) -> None:
        raise NotImplementedError

    def shutdown(self) -> None:
        """Flush / close the sink."""
