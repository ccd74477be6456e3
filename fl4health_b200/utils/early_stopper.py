"""In-round early stopping on validation loss (parity: ``fl4health/utils/early_stopper.py:14-98``)."""

from __future__ import annotations

from logging import INFO, WARNING
from pathlib import Path
from typing import TYPE_CHECKING

from fl4health_b200.checkpointing.state_checkpointer import ClientStateCheckpointer
from fl4health_b200.common.logger import log
from fl4health_b200.utils.logging import LoggingMode

if TYPE_CHECKING:
    from fl4health_b200.clients.basic_client import BasicClient


class EarlyStopper:
    def __init__(
        self,
        client: BasicClient,
        patience: int | None = 1,
        interval_steps: int = 5,
        snapshot_dir: Path | None = None,
        train_loop_checkpoint_dir: Path | None = None,
    ) -> None:
        """``train_loop_checkpoint_dir`` is the reference's name for ``snapshot_dir`` (``early_stopper.py:14-60``)."""
        if snapshot_dir is None:
            snapshot_dir = train_loop_checkpoint_dir
        self.client = client
        self.patience = patience
        self.count_down = patience
        self.interval_steps = interval_steps
        self.best_score: float | None = None
        self.snapshot_ckpt: dict = {}
        checkpoint_name = f"temp_{client.client_name}.pt"
        self.state_checkpointer = ClientStateCheckpointer(
            checkpoint_dir=snapshot_dir if snapshot_dir is not None else Path("."), checkpoint_name=checkpoint_name
        )
        if snapshot_dir is None:
            log(WARNING, "EarlyStopper snapshots go to the current directory because snapshot_dir is None")

    def load_snapshot(self, attributes: list[str] | None = None) -> None:
        self.state_checkpointer.maybe_load_client_state(self.client, attributes)

    def should_stop(self, steps: int) -> bool:
        if steps % self.interval_steps != 0:
            return False
        val_loss, _ = self.client._fully_validate_or_test(
            loader=self.client.val_loader,
            loss_meter=self.client.val_loss_meter,
            metric_manager=self.client.val_metric_manager,
            logging_mode=LoggingMode.EARLY_STOP_VALIDATION,
            include_losses_in_metrics=False,
        )
        self.client.model.train()
        if val_loss is None:
            return False
        if self.best_score is None or val_loss < self.best_score:
            self.best_score = val_loss
            self.count_down = self.patience
            self.state_checkpointer.save_client_state(self.client)
            return False
        if self.count_down is not None:
            self.count_down -= 1
            if self.count_down <= 0:
                log(INFO, "Early stopping patience exhausted.")
                return True
        return False
