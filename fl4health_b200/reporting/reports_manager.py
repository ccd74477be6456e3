"""Fan-out of report calls to every configured reporter (parity: ``fl4health/reporting/reports_manager.py:7-29``)."""

from __future__ import annotations

from collections.abc import Sequence
from typing import Any

from fl4health_b200.reporting.base_reporter import BaseReporter


class ReportsManager:
    def __init__(self, reporters: Sequence[BaseReporter] | None = None) -> None:
        self.reporters: list[BaseReporter] = list(reporters) if reporters is not None else []

    def initialize(self, **kwargs: Any) -> None:
        for reporter in self.reporters:
            reporter.initialize(**kwargs)

    def report(self, data: dict, round: int | None = None, epoch: int | None = None, step: int | None = None) -> None:
        for reporter in self.reporters:
            reporter.report(data, round, epoch, step)

    def shutdown(self) -> None:
        for reporter in self.reporters:
            reporter.shutdown()
