"""What one server instruction asks of a client, parsed once, plus the wall-clock bookkeeping of a round's reports.

The config keys and the report keys are the reference's wire contract (``fl4health/clients/basic_client.py:255-290`` for
the keys, ``:350-375, 415-430`` for the payloads); the parsing / stopwatch objects are this engine's."""

from __future__ import annotations

import datetime
from dataclasses import dataclass
from typing import Any

from fl4health_b200.common.typing import Config
from fl4health_b200.utils.client import set_pack_losses_with_val_metrics
from fl4health_b200.utils.config import narrow_dict_type


@dataclass(frozen=True)
class RoundPlan:
    """Local work requested for one round: exactly one of ``epochs`` / ``steps`` is set."""

    server_round: int
    epochs: int | None
    steps: int | None
    evaluate_after_fit: bool
    pack_losses_with_val_metrics: bool

    @classmethod
    def from_config(cls, config: Config) -> "RoundPlan":
        server_round = narrow_dict_type(config, "current_server_round", int)
        asked = [key for key in ("local_epochs", "local_steps") if key in config]
        if len(asked) == 2:
            raise ValueError("Config cannot contain both local_epochs and local_steps. Please specify only one.")
        if not asked:
            raise ValueError("Must specify either local_epochs or local_steps in the Config.")
        amount = narrow_dict_type(config, asked[0], int)
        return cls(
            server_round=server_round,
            epochs=amount if asked[0] == "local_epochs" else None,
            steps=amount if asked[0] == "local_steps" else None,
            evaluate_after_fit=bool(config.get("evaluate_after_fit", False)),
            pack_losses_with_val_metrics=set_pack_losses_with_val_metrics(config),
        )

    def as_tuple(self) -> tuple[int | None, int | None, int, bool, bool]:
        """The reference's ``process_config`` return order."""
        return self.epochs, self.steps, self.server_round, self.evaluate_after_fit, self.pack_losses_with_val_metrics


class Stopwatch:
    """Named wall-clock marks; ``span`` renders the (start, end, whole seconds elapsed) triple the reports carry."""

    def __init__(self) -> None:
        self.marks: dict[str, datetime.datetime] = {}

    def mark(self, name: str) -> datetime.datetime:
        self.marks[name] = datetime.datetime.now()
        return self.marks[name]

    def span(self, prefix: str, start: str, end: str) -> dict[str, Any]:
        begin, finish = self.marks[start], self.marks[end]
        return {
            f"{prefix}_start": str(begin),
            f"{prefix}_time_elapsed": round((finish - begin).total_seconds()),
            f"{prefix}_end": str(finish),
        }
