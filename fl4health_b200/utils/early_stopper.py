"""In-round early stopping on validation loss (parity: ``fl4health/utils/early_stopper.py:14-98``)."""

from __future__ import annotations

from logging import INFO, WARNING
from pathlib import Path
from typing import TYPE_CHECKING

from fl4health_b200.checkpointing.state_checkpointer import ClientStateCheckpointer
from fl4health_b200.common.logger import log
from fl4health_b200.engine.companions import set_phase
from fl4health_b200.utils.logging import LoggingMode

if TYPE_CHECKING:
    from fl4health_b200.clients.basic_client import BasicClient


class EarlyStopper:
    """Every ``interval_steps`` steps: score the model on the client's validation loader; a new best is snapshotted (whole
    client training state, through ``ClientStateCheckpointer``); ``patience`` non-improving checks end local training
    (``patience=None``: never stop, only keep the best snapshot for ``load_snapshot``)."""

    def __init__(
        self,
        client: BasicClient,
        patience: int | None = 1,
        interval_steps: int = 5,
        snapshot_dir: Path | None = None,
        train_loop_checkpoint_dir: Path | None = None,
    ) -> None:
        """``train_loop_checkpoint_dir`` is the reference's name for ``snapshot_dir`` (``early_stopper.py:14-60``)."""
        directory = snapshot_dir if snapshot_dir is not None else train_loop_checkpoint_dir
        if directory is None:
            log(WARNING, "EarlyStopper snapshots go to the current directory because snapshot_dir is None")
            directory = Path(".")
        self.client, self.patience, self.interval_steps = client, patience, interval_steps
        self.count_down = patience
        self.best_score: float | None = None
        self.snapshot_ckpt: dict = {}
        self.state_checkpointer = ClientStateCheckpointer(checkpoint_dir=directory, checkpoint_name=f"temp_{client.client_name}.pt")

    def load_snapshot(self, attributes: list[str] | None = None) -> None:
        self.state_checkpointer.maybe_load_client_state(self.client, attributes)

    def _score(self) -> float | None:
        client = self.client
        loss, _ = client._fully_validate_or_test(client.val_loader, client.val_loss_meter, client.val_metric_manager,
                                                 LoggingMode.EARLY_STOP_VALIDATION, include_losses_in_metrics=False)
        set_phase(client, training=True)  # scoring flipped every network to eval
        return loss

    def should_stop(self, steps: int) -> bool:
        if steps % self.interval_steps:
            return False
        score = self._score()
        if score is None:
            return False
        improved = self.best_score is None or score < self.best_score
        if improved:
            self.best_score, self.count_down = score, self.patience
            self.state_checkpointer.save_client_state(self.client)
            return False
        if self.count_down is None:
            return False
        self.count_down -= 1
        exhausted = self.count_down <= 0
        if exhausted:
            log(INFO, "Early stopping patience exhausted.")
        return exhausted
