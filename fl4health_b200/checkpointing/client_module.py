"""Client-side checkpoint + state bundle (parity: ``fl4health/checkpointing/client_module.py:23-164``)."""

from __future__ import annotations

from collections.abc import Sequence
from enum import Enum
from logging import INFO
from typing import TYPE_CHECKING

from torch import nn

from fl4health_b200.checkpointing.checkpointer import TorchModuleCheckpointer
from fl4health_b200.checkpointing.state_checkpointer import ClientStateCheckpointer
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Scalar

if TYPE_CHECKING:
    from fl4health_b200.clients.basic_client import BasicClient

ModelCheckpointers = TorchModuleCheckpointer | Sequence[TorchModuleCheckpointer] | None


class CheckpointMode(Enum):
    PRE_AGGREGATION = "pre_aggregation"
    POST_AGGREGATION = "post_aggregation"


def _as_list(checkpointers: ModelCheckpointers) -> list[TorchModuleCheckpointer] | None:
    if checkpointers is None:
        return None
    if isinstance(checkpointers, TorchModuleCheckpointer):
        return [checkpointers]
    return list(checkpointers)


class ClientCheckpointAndStateModule:
    def __init__(
        self,
        pre_aggregation: ModelCheckpointers = None,
        post_aggregation: ModelCheckpointers = None,
        state_checkpointer: ClientStateCheckpointer | None = None,
    ) -> None:
        self.pre_aggregation = _as_list(pre_aggregation)
        self.post_aggregation = _as_list(post_aggregation)
        self._check_if_shared_checkpoint_names()
        self.state_checkpointer = state_checkpointer

    def _check_if_shared_checkpoint_names(self) -> None:
        paths = [c.checkpoint_path for c in (self.pre_aggregation or [])]
        paths += [c.checkpoint_path for c in (self.post_aggregation or [])]
        if len(set(paths)) != len(paths):
            listing = "\n".join(paths)
            raise ValueError(
                "The paths of all of your checkpointers should be unique otherwise overwrites are possible and data "
                f"will be lost. The current paths are:\n{listing}"
            )

    def maybe_checkpoint(
        self, model: nn.Module, loss: float, metrics: dict[str, Scalar], mode: CheckpointMode
    ) -> None:
        if mode == CheckpointMode.PRE_AGGREGATION:
            chosen = self.pre_aggregation
        elif mode == CheckpointMode.POST_AGGREGATION:
            chosen = self.post_aggregation
        else:
            raise ValueError(f"Unrecognized mode for checkpointing: {mode}")
        if chosen is None:
            log(INFO, f"No {mode.value} checkpoint specified. Skipping.")
            return
        for checkpointer in chosen:
            checkpointer.maybe_checkpoint(model, loss, metrics)

    def save_state(self, client: BasicClient) -> None:
        if self.state_checkpointer is None:
            raise ValueError("Attempting to save state but no state checkpointer is specified")
        self.state_checkpointer.save_client_state(client)

    def maybe_load_state(self, client: BasicClient, attributes: list[str] | None = None) -> bool:
        if self.state_checkpointer is None:
            raise ValueError("Attempting to load state, but no state checkpointer is specified")
        return self.state_checkpointer.maybe_load_client_state(client, attributes)
