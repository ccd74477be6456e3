"""The federated round loop as an explicit phase machine.

    BOOTSTRAP -> ( FIT -> CENTRAL_EVAL -> FEDERATED_EVAL -> CLOSE )  x num_rounds

``FlServer`` (in-process simulation, SPMD one-client-per-GPU, clients-per-rank) owns the *content* of each phase
(``fit_round`` / ``evaluate_round`` / strategy calls); this module owns the *sequencing*: which phase runs when, what is
written to the ``History`` after it, where resumable state is loaded and saved, where observers (``round_end_hooks``) and
fault injection fire.  The reference inlines this sequencing twice (Flower's ``Server.fit`` and
``fl4health/servers/base_server.py:143-230``, the per-round-checkpointing copy of it).
"""

from __future__ import annotations

import datetime
import os
from collections.abc import Callable
from enum import Enum
from logging import INFO
from typing import Any

from fl4health_b200.common.history import History
from fl4health_b200.common.logger import log


class Phase(Enum):
    BOOTSTRAP = "bootstrap"
    FIT = "fit"
    CENTRAL_EVAL = "central_eval"
    FEDERATED_EVAL = "federated_eval"
    CLOSE = "close"


ROUND_PHASES = (Phase.FIT, Phase.CENTRAL_EVAL, Phase.FEDERATED_EVAL, Phase.CLOSE)


class RoundMachine:
    def __init__(self, server: Any, num_rounds: int, timeout: float | None, resumable: bool) -> None:
        self.server, self.num_rounds, self.timeout, self.resumable = server, num_rounds, timeout, resumable
        self.started_at: datetime.datetime | None = None
        self.handlers: dict[Phase, Callable[[int], None]] = {
            Phase.FIT: self._fit,
            Phase.CENTRAL_EVAL: self._central_eval,
            Phase.FEDERATED_EVAL: self._federated_eval,
            Phase.CLOSE: self._close,
        }

    # -- BOOTSTRAP: global parameters, fresh or restored history ----------------------------------------------
    def bootstrap(self) -> None:
        server = self.server
        log(INFO, "Initializing server state and global parameters")
        server.parameters = server._get_initial_parameters(server_round=0, timeout=self.timeout)
        server.history, server.current_round = History(), 1
        if self.resumable:
            restored = server._load_server_state()
            log(INFO, "Server state checkpoint successfully loaded." if restored
                else "No server state checkpoint found. Starting from scratch.")
        if server.current_round == 1:  # nothing restored: score the initial model centrally, if the strategy can
            log(INFO, "Evaluating initial parameters")
            self._record_central(0, announce="initial parameters (loss, other metrics): %s, %s")
            log(INFO, "FL starting")

    def _record_central(self, server_round: int, announce: str | None = None) -> tuple[float, dict] | None:
        outcome = self.server.strategy.evaluate(server_round, parameters=self.server.parameters)
        if outcome is None:
            return None
        loss, metrics = outcome
        if announce:
            log(INFO, announce, loss, metrics)
        self.server.history.add_loss_centralized(server_round=server_round, loss=loss)
        self.server.history.add_metrics_centralized(server_round=server_round, metrics=metrics)
        return loss, metrics

    # -- round phases -----------------------------------------------------------------------------------------
    def _fit(self, server_round: int) -> None:
        outcome = self.server.fit_round(server_round=server_round, timeout=self.timeout)
        if not outcome:
            return
        new_parameters, fit_metrics, _ = outcome
        if new_parameters:
            self.server.parameters = new_parameters
        self.server.history.add_metrics_distributed_fit(server_round=server_round, metrics=fit_metrics)

    def _central_eval(self, server_round: int) -> None:
        scored = self._record_central(server_round)
        if scored is not None:
            assert self.started_at is not None
            log(INFO, "fit progress: (%s, %s, %s, %s)", server_round, scored[0], scored[1],
                (datetime.datetime.now() - self.started_at).total_seconds())

    def _federated_eval(self, server_round: int) -> None:
        outcome = self.server.evaluate_round(server_round=server_round, timeout=self.timeout)
        if outcome and outcome[0] is not None:
            self.server.history.add_loss_distributed(server_round=server_round, loss=outcome[0])
            self.server.history.add_metrics_distributed(server_round=server_round, metrics=outcome[1])

    def _close(self, server_round: int) -> None:
        server = self.server
        for observer in server.round_end_hooks:
            observer(server_round)
        server.current_round = server_round + 1
        if self.resumable:
            server._save_server_state()
        inject_fault_after(server_round)

    # -- driver -----------------------------------------------------------------------------------------------
    def run(self) -> tuple[History, float]:
        self.bootstrap()
        self.started_at = datetime.datetime.now()
        while self.server.current_round <= self.num_rounds:
            server_round = self.server.current_round
            log(INFO, "[ROUND %s]", server_round)
            for phase in ROUND_PHASES:
                self.handlers[phase](server_round)
        elapsed = datetime.datetime.now() - self.started_at
        log(INFO, "FL finished in %s", str(elapsed))
        return self.server.history, elapsed.total_seconds()


def inject_fault_after(finished_round: int) -> None:
    """Fault injection for resume tests: ``FL4H_FAULT_AFTER_ROUND=r`` aborts the process after round r finished (and its
    state was saved), imitating a pre-emption."""
    target = os.environ.get("FL4H_FAULT_AFTER_ROUND")
    if target is not None and int(target) == finished_round:
        raise SystemExit(f"fault injected after round {target}")
