"""FLASH client (parity: ``fl4health/clients/flash_client.py:18-176``): epoch-level early stop — training ends once the
validation-loss improvement drops below ``gamma / (epoch + 1)``.  ``gamma`` comes from the server config.

The reference re-implements the whole epoch loop to slot the rule in; here the shared loop driver
(``engine/local_loop.run_training``) asks ``stop_after_epoch`` at every epoch boundary, so the client is only the rule."""

from __future__ import annotations

from logging import INFO
from typing import Any

from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config, Scalar
from fl4health_b200.utils.config import narrow_dict_type


class FlashClient(BasicClient):
    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self.gamma: float | None = None
        self._previous_validation_loss = float("inf")

    def setup_client(self, config: Config) -> None:
        super().setup_client(config)
        if "gamma" not in config:
            log(INFO, "Gamma not present in config. Early stopping is disabled.")
            return
        self.gamma = narrow_dict_type(config, "gamma", float)

    def process_config(self, config: Config) -> tuple[int | None, int | None, int, bool, bool]:
        plan = super().process_config(config)
        if plan[1] is not None and self.gamma is not None:
            raise ValueError("Training by steps is not applicable for FLASH clients with gamma defined (epochs only).")
        return plan

    def train_by_epochs(self, epochs: int, current_round: int | None = None) -> tuple[dict[str, float], dict[str, Scalar]]:
        self._previous_validation_loss = float("inf")  # the rule compares epochs of ONE round
        return super().train_by_epochs(epochs, current_round)

    def stop_after_epoch(self, epoch: int) -> bool:
        if self.gamma is None:
            return False
        current, _ = self.validate()
        improvement = self._previous_validation_loss - current
        if improvement < self.gamma / (epoch + 1):
            log(INFO, f"Early stopping at epoch {epoch} with loss change {abs(improvement)} and gamma {self.gamma}")
            return True
        self._previous_validation_loss = current
        return False
