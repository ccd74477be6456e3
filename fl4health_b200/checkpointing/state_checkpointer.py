"""Training-state ("restart") checkpoints.

Parity: ``fl4health/checkpointing/state_checkpointer.py:41-585``: an attribute-name -> (snapshotter, type) map is
applied to the live client/server object; the result is ``torch.save``d to ``client_{name}_state.pt`` /
``server_{name}_state.pt``.  Scalars are wrapped as ``{"None": value}`` so one code path serves scalars and dicts.
"""

from __future__ import annotations

import os
from abc import ABC
from collections.abc import Iterator
from contextlib import contextmanager
from enum import Enum
from logging import ERROR, INFO, WARNING
from pathlib import Path
from typing import TYPE_CHECKING, Any

import torch
from torch import nn
from torch.optim import Optimizer
from torch.optim.lr_scheduler import LRScheduler

from fl4health_b200.common.history import History
from fl4health_b200.common.logger import log
from fl4health_b200.metrics.metric_managers import MetricManager
from fl4health_b200.reporting.reports_manager import ReportsManager
from fl4health_b200.utils.losses import LossMeter
from fl4health_b200.utils.snapshotter import (
    AbstractSnapshotter,
    BytesSnapshotter,
    EnumSnapshotter,
    HistorySnapshotter,
    LRSchedulerSnapshotter,
    OptimizerSnapshotter,
    SerializableObjectSnapshotter,
    SingletonSnapshotter,
    StringSnapshotter,
    TorchModuleSnapshotter,
)

if TYPE_CHECKING:
    from fl4health_b200.clients.basic_client import BasicClient
    from fl4health_b200.servers.base_server import FlServer

SnapshotSpec = dict[str, tuple[AbstractSnapshotter, Any]]

_SCALAR_KEY = "None"  # a lone value is stored as {"None": value} so snapshotters only ever see dictionaries


class StateCheckpointer(ABC):
    """Applies a table ``attribute name -> (snapshotter, value type)`` to a live object (the *host*: a client or a
    server) and persists the result as one ``torch.save`` file.  Subclasses only say how the host is bound and what the
    default file is called; attributes that do not live on the host itself (the server's model) are kept in
    ``self._detached`` while a save / load is in progress."""

    host_kind = "object"

    def __init__(self, checkpoint_dir: Path, checkpoint_name: str | None, snapshot_attrs: SnapshotSpec) -> None:
        self.checkpoint_dir, self.checkpoint_name = checkpoint_dir, checkpoint_name
        self.checkpoint_path: str | None = None if checkpoint_name is None else os.path.join(checkpoint_dir, checkpoint_name)
        self.snapshot_attrs = snapshot_attrs
        self.snapshot_ckpt: dict[str, Any] = {}
        self._host: Any = None
        self._detached: dict[str, Any] = {}

    # ---- where the file lives --------------------------------------------------------------------------------
    def set_checkpoint_path(self, checkpoint_dir: Path, checkpoint_name: str) -> None:
        self.checkpoint_dir, self.checkpoint_name = checkpoint_dir, checkpoint_name
        self.checkpoint_path = os.path.join(checkpoint_dir, checkpoint_name)

    def checkpoint_exists(self) -> bool:
        assert self.checkpoint_path is not None, "A checkpoint_path should be set but is not"
        return os.path.exists(self.checkpoint_path)

    def _default_file_name(self) -> str:
        raise NotImplementedError

    def maybe_set_default_checkpoint_name(self) -> None:
        assert self._host is not None, f"Attempting to save {self.host_kind} state but {self.host_kind} is None"
        if self.checkpoint_name is None:
            self.set_checkpoint_path(self.checkpoint_dir, self._default_file_name())

    @contextmanager
    def _bound(self, host: Any, **detached: Any) -> Iterator[None]:
        """Bind the host (and detached attributes) for the duration of one save / load, whatever happens inside."""
        self._host, self._detached = host, dict(detached)
        try:
            self.maybe_set_default_checkpoint_name()
            yield
        finally:
            self._host = None

    # ---- file I/O --------------------------------------------------------------------------------------------
    def save_checkpoint(self, checkpoint_dict: dict[str, Any]) -> None:
        assert self.checkpoint_path is not None, "Checkpoint path is not set but save_checkpoint has been called."
        scratch = f"{self.checkpoint_path}.tmp"  # write-then-rename: a pre-emption mid-write keeps the previous state intact
        try:
            torch.save(checkpoint_dict, scratch)
            os.replace(scratch, self.checkpoint_path)
        except Exception as exc:
            log(ERROR, f"Encountered the following error while saving the checkpoint: {exc}")
            raise

    def load_checkpoint(self) -> dict[str, Any]:
        assert self.checkpoint_path is not None, "Checkpoint path is not set but load_checkpoint has been called."
        assert self.checkpoint_exists(), f"Could not verify existence of checkpoint file at {self.checkpoint_path}"
        log(INFO, f"Loading state from checkpoint at {self.checkpoint_path}")
        return torch.load(self.checkpoint_path, weights_only=False)

    # ---- the attribute table ---------------------------------------------------------------------------------
    def add_to_snapshot_attr(self, name: str, snapshotter: AbstractSnapshotter, input_type: Any) -> None:
        self.snapshot_attrs[name] = (snapshotter, input_type)

    def delete_from_snapshot_attr(self, name: str) -> None:
        del self.snapshot_attrs[name]

    def _dict_wrap_attr(self, name: str, expected_type: Any) -> dict[str, Any]:
        """The attribute as ``{key: value}``: a dictionary of ``expected_type`` values as is, a lone value under "None"."""
        value = self.get_attribute(name)
        if isinstance(value, expected_type):
            return {_SCALAR_KEY: value}
        if not isinstance(value, dict):
            raise ValueError(f"Incompatible type of attribute {type(value)}, expected {expected_type}")
        offending = next((key for key, item in value.items() if not isinstance(item, expected_type)), None)
        if offending is not None:
            raise ValueError(f"Incompatible type of attribute {type(value)} for key {offending}")
        return value

    def save_state(self) -> None:
        self.snapshot_ckpt = {name: snapshotter.save_attribute(self._dict_wrap_attr(name, kind))
                              for name, (snapshotter, kind) in self.snapshot_attrs.items()}
        log(INFO, f"Saving the state to checkpoint at {self.checkpoint_path}")
        self.save_checkpoint(self.snapshot_ckpt)
        self.snapshot_ckpt = {}

    def load_state(self, attributes: list[str] | None = None) -> None:
        assert self.checkpoint_exists(), f"No state checkpoint to load. {self.checkpoint_path} does not exist"
        wanted = list(self.snapshot_attrs) if attributes is None else attributes
        if not wanted:
            log(WARNING, "self.snapshot_attrs is empty, which may be undesired behavior.")
        self.snapshot_ckpt = self.load_checkpoint()
        for name in wanted:
            snapshotter, kind = self.snapshot_attrs[name]
            live = self._dict_wrap_attr(name, kind)
            snapshotter.load_attribute(self.snapshot_ckpt[name], live)
            self.set_attribute(name, live[_SCALAR_KEY] if set(live) == {_SCALAR_KEY} else live)
        log(INFO, f"Loaded the checkpointed state from {self.checkpoint_path}")
        self.snapshot_ckpt = {}

    def _maybe_load_bound(self, attributes: list[str] | None) -> bool:
        if not self.checkpoint_exists():
            log(INFO, f"No state checkpoint found at: {self.checkpoint_path}")
            return False
        self.load_state(attributes)
        log(INFO, f"State checkpoint successfully loaded from: {self.checkpoint_path}")
        return True

    def get_attribute(self, name: str) -> Any:
        assert self._host is not None, f"{self.host_kind.capitalize()} is not set."
        return self._detached[name] if name in self._detached else getattr(self._host, name)

    def set_attribute(self, name: str, value: Any) -> None:
        assert self._host is not None, f"{self.host_kind.capitalize()} is not set."
        if name in self._detached:
            self._detached[name] = value
        else:
            setattr(self._host, name, value)


def default_client_snapshot_attrs() -> SnapshotSpec:
    whole_object = SerializableObjectSnapshotter
    return {
        "model": (TorchModuleSnapshotter(), nn.Module),
        "optimizers": (OptimizerSnapshotter(), Optimizer),
        "lr_schedulers": (LRSchedulerSnapshotter(), LRScheduler),
        **{counter: (SingletonSnapshotter(), int) for counter in ("total_steps", "total_epochs")},
        "reports_manager": (whole_object(), ReportsManager),
        "train_loss_meter": (whole_object(), LossMeter),
        "train_metric_manager": (whole_object(), MetricManager),
    }


class ClientStateCheckpointer(StateCheckpointer):
    host_kind = "client"

    def __init__(
        self, checkpoint_dir: Path, checkpoint_name: str | None = None, snapshot_attrs: SnapshotSpec | None = None
    ) -> None:
        super().__init__(checkpoint_dir, checkpoint_name, snapshot_attrs or default_client_snapshot_attrs())

    @property
    def client(self) -> BasicClient | None:
        return self._host

    def _default_file_name(self) -> str:
        return f"client_{self._host.client_name}_state.pt"

    def save_client_state(self, client: BasicClient) -> None:
        with self._bound(client):
            self.save_state()

    def maybe_load_client_state(self, client: BasicClient, attributes: list[str] | None = None) -> bool:
        with self._bound(client):
            return self._maybe_load_bound(attributes)


def default_server_snapshot_attrs() -> SnapshotSpec:
    return {
        "model": (TorchModuleSnapshotter(), nn.Module),
        "current_round": (SingletonSnapshotter(), int),
        "reports_manager": (SerializableObjectSnapshotter(), ReportsManager),
        "server_name": (StringSnapshotter(), str),
        "history": (HistorySnapshotter(), History),
    }


class ServerStateCheckpointer(StateCheckpointer):
    """The server's model is not an attribute of the server (the strategy holds parameters, the checkpoint module holds
    the architecture): it travels as a detached attribute."""

    host_kind = "server"

    def __init__(
        self, checkpoint_dir: Path, checkpoint_name: str | None = None, snapshot_attrs: SnapshotSpec | None = None
    ) -> None:
        super().__init__(checkpoint_dir, checkpoint_name, snapshot_attrs or default_server_snapshot_attrs())

    @property
    def server(self) -> FlServer | None:
        return self._host

    @property
    def server_model(self) -> nn.Module | None:
        return self._detached.get("model")

    def _default_file_name(self) -> str:
        return f"server_{self._host.server_name}_state.pt"

    def save_server_state(self, server: FlServer, model: nn.Module) -> None:
        with self._bound(server, model=model):
            self.save_state()
        self._detached = {}

    def maybe_load_server_state(
        self, server: FlServer, model: nn.Module, attributes: list[str] | None = None
    ) -> nn.Module | None:
        with self._bound(server, model=model):
            restored = self._maybe_load_bound(attributes)
        loaded_model, self._detached = self._detached.get("model"), {}
        return loaded_model if restored else None


class NnUnetServerStateCheckpointer(ServerStateCheckpointer):
    def __init__(self, checkpoint_dir: Path, checkpoint_name: str | None = None) -> None:
        plan_negotiation: SnapshotSpec = {
            "nnunet_plans_bytes": (BytesSnapshotter(), bytes),
            "num_segmentation_heads": (SingletonSnapshotter(), int),
            "num_input_channels": (SingletonSnapshotter(), int),
            "global_deep_supervision": (EnumSnapshotter(), bool),
            "nnunet_config": (EnumSnapshotter(), Enum),
        }
        super().__init__(checkpoint_dir, checkpoint_name, {**default_server_snapshot_attrs(), **plan_negotiation})
