"""Server-side controller of a penalty weight driven by the aggregate training loss.

FedProx's adaptive ``mu`` (and Ditto / MR-MTL's ``lambda``) follow one rule: every round the loss fails to increase counts
towards ``patience``; at ``patience`` consecutive such rounds the weight shrinks by ``delta`` (never below zero) and the
count restarts; any increase of the loss grows the weight by ``delta`` at once.  The reference inlines the rule in
``fl4health/strategies/fedavg_with_adaptive_constraint.py:190-232``; two strategies use it here."""

from __future__ import annotations

from dataclasses import dataclass, field
from logging import INFO

from fl4health_b200.common.logger import log


@dataclass
class LossDrivenWeight:
    value: float
    adaptive: bool = False
    delta: float = 0.1
    patience: int = 5
    calm_rounds: int = field(default=0, init=False)
    last_loss: float = field(default=float("inf"), init=False)

    def observe(self, loss: float) -> float:
        """Feed this round's aggregate training loss; returns the (possibly updated) weight."""
        if self.adaptive:
            if loss > self.last_loss:
                self.value += self.delta
                self.calm_rounds = 0
                log(INFO, f"Aggregate training loss increased this round: Current loss {loss}, Previous loss: {self.last_loss}")
                log(INFO, f"Constraint weight is increased by {self.delta} to {self.value}")
            else:
                self.calm_rounds += 1
                if self.calm_rounds == self.patience:
                    self.value = max(0.0, self.value - self.delta)
                    self.calm_rounds = 0
                    log(INFO, f"Aggregate training loss has dropped {self.patience} rounds in a row")
                    log(INFO, f"Constraint weight is decreased to {self.value}")
        self.last_loss = loss
        return self.value
