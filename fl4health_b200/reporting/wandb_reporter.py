"""Weights & Biases reporter (parity: ``fl4health/reporting/wandb_reporter.py:21-248``).

``wandb`` is an optional dependency; constructing the reporter without it raises a clear error.  Data is logged
against a round / epoch / step x-axis selected by ``wandb_step_type``.
"""

from __future__ import annotations

from collections.abc import Iterable
from enum import Enum
from pathlib import Path
from typing import Any

from fl4health_b200.reporting.base_reporter import BaseReporter


class WandBStepType(Enum):
    ROUND = "round"
    EPOCH = "epoch"
    STEP = "step"


class WandBReporter(BaseReporter):
    def __init__(
        self,
        wandb_step_type: WandBStepType | str = WandBStepType.ROUND,
        project: str | None = None,
        entity: str | None = None,
        config: dict | str | None = None,
        group: str | None = None,
        job_type: str | None = None,
        tags: list[str] | None = None,
        name: str | None = None,
        id: str | None = None,
        resume: str = "allow",
        **kwargs: Any,
    ) -> None:
        try:
            import wandb  # type: ignore[import-not-found]
        except ImportError as exc:  # pragma: no cover - optional dependency
            raise ImportError("WandBReporter requires the optional `wandb` package") from exc
        self._wandb = wandb
        self.wandb_init_kwargs = dict(kwargs)
        self.wandb_step_type = WandBStepType(wandb_step_type)
        self.project, self.entity, self.config, self.group = project, entity, config, group
        self.job_type, self.tags, self.name, self.id, self.resume = job_type, tags, name, id, resume
        self.initialized = False
        self.timestamp = None
        self.run_started = False
        self.run: Any = None
        self.current_x = 0

    def initialize(self, **kwargs: Any) -> None:
        if self.name is None:
            self.name = kwargs.get("name")
        if self.id is None:
            self.id = kwargs.get("id")
        self.initialized = True

    def define_metrics(self) -> None:
        self.run.define_metric("fit_step")
        self.run.define_metric("fit_epoch")
        self.run.define_metric("round")
        self.run.define_metric("round_start")
        self.run.define_metric("round_end")
        self.run.define_metric("fit_round_time_elapsed")
        self.run.define_metric("eval_round_time_elapsed")

    def start_run(self, wandb_init_kwargs: dict[str, Any]) -> None:
        if not self.initialized:
            self.initialize()
        self.run = self._wandb.init(
            project=self.project, entity=self.entity, config=self.config, group=self.group, job_type=self.job_type,
            tags=self.tags, name=self.name, id=self.id, resume=self.resume, **wandb_init_kwargs,
        )
        self.run_id = self.run._run_id
        self.run_started = True
        self.define_metrics()

    def get_wandb_timestep(self, round: int | None, epoch: int | None, step: int | None) -> int | None:
        if self.wandb_step_type == WandBStepType.ROUND and round is not None:
            return round
        if self.wandb_step_type == WandBStepType.EPOCH and epoch is not None:
            return epoch
        if self.wandb_step_type == WandBStepType.STEP and step is not None:
            return step
        return None

    def report(
        self, data: dict[str, Any], round: int | None = None, epoch: int | None = None, step: int | None = None
    ) -> None:
        if not self.run_started:
            self.start_run(self.wandb_init_kwargs)
        if self.wandb_step_type == WandBStepType.ROUND and (epoch is not None or step is not None):
            return
        if self.wandb_step_type == WandBStepType.EPOCH and step is not None:
            return
        flat: dict[str, Any] = {}
        for key, value in data.items():
            if isinstance(value, dict):
                flat.update({f"{key}/{k}" if not isinstance(k, str) else k: v for k, v in value.items()})
            elif not isinstance(value, Iterable) or isinstance(value, str):
                flat[key] = value
        self.run.log(flat)

    def shutdown(self) -> None:
        if self.run is not None:
            self.run.finish()


def _unused(_: Path) -> None:  # keeps Path import meaningful for type checkers
    return None
