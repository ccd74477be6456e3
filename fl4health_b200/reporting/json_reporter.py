"""File reporters (parity: ``fl4health/reporting/json_reporter.py:12-97``).

``JsonReporter`` writes ``{output_folder}/{run_id}.json`` with the schema ``{..., "rounds": {r: {...}}}`` that the
reference's smoke tests diff against golden files.  Per-step / per-epoch calls are ignored, as in the reference.
"""

from __future__ import annotations

import json
import uuid
from logging import INFO
from pathlib import Path
from typing import Any

from fl4health_b200.common.logger import log
from fl4health_b200.reporting.base_reporter import BaseReporter


class FileReporter(BaseReporter):
    def __init__(self, run_id: str | None = None, output_folder: str | Path = Path("metrics")) -> None:
        self.run_id = run_id
        self.output_folder = Path(output_folder)
        self.metrics: dict[str, Any] = {}
        self.initialized = False
        self.output_folder.mkdir(exist_ok=True, parents=True)

    def initialize(self, **kwargs: Any) -> None:
        if self.run_id is None:
            self.run_id = kwargs.get("id") or str(uuid.uuid4())
        self.initialized = True

    def report(
        self, data: dict[str, Any], round: int | None = None, epoch: int | None = None, step: int | None = None
    ) -> None:
        if not self.initialized:
            self.initialize()
        if round is None:
            self.metrics.update(data)
        elif epoch is None and step is None:
            self.metrics.setdefault("rounds", {}).setdefault(round, {}).update(data)

    def dump(self) -> None:
        raise NotImplementedError

    def shutdown(self) -> None:
        self.dump()


def _jsonable(value: Any) -> Any:
    if isinstance(value, dict):
        return {str(k): _jsonable(v) for k, v in value.items()}
    if isinstance(value, (list, tuple)):
        return [_jsonable(v) for v in value]
    if isinstance(value, bytes):
        return value.decode(errors="replace")
    if hasattr(value, "item") and callable(value.item):
        try:
            return value.item()
        except Exception:  # noqa: BLE001 - multi-element tensors/arrays
            return value.tolist() if hasattr(value, "tolist") else str(value)
    return value


class JsonReporter(FileReporter):
    def dump(self) -> None:
        assert isinstance(self.run_id, str)
        output_file_path = Path(self.output_folder, self.run_id).with_suffix(".json")
        log(INFO, f"Dumping metrics to {output_file_path}")
        with open(output_file_path, "w") as handle:
            json.dump(_jsonable(self.metrics), handle, indent=4)
