"""Training-state ("restart") checkpoints.

Parity: ``fl4health/checkpointing/state_checkpointer.py:41-585``: an attribute-name -> (snapshotter, type) map is
applied to the live client/server object; the result is ``torch.save``d to ``client_{name}_state.pt`` /
``server_{name}_state.pt``.  Scalars are wrapped as ``{"None": value}`` so one code path serves scalars and dicts.
"""

from __future__ import annotations

import os
from abc import ABC, abstractmethod
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


class StateCheckpointer(ABC):
    def __init__(self, checkpoint_dir: Path, checkpoint_name: str | None, snapshot_attrs: SnapshotSpec) -> None:
        self.checkpoint_dir = checkpoint_dir
        self.checkpoint_name = checkpoint_name
        self.checkpoint_path: str | None = (
            os.path.join(checkpoint_dir, checkpoint_name) if checkpoint_name is not None else None
        )
        self.snapshot_attrs = snapshot_attrs
        self.snapshot_ckpt: dict[str, Any] = {}

    def set_checkpoint_path(self, checkpoint_dir: Path, checkpoint_name: str) -> None:
        self.checkpoint_dir, self.checkpoint_name = checkpoint_dir, checkpoint_name
        self.checkpoint_path = os.path.join(checkpoint_dir, checkpoint_name)

    def checkpoint_exists(self) -> bool:
        assert self.checkpoint_path is not None, "A checkpoint_path should be set but is not"
        return os.path.exists(self.checkpoint_path)

    def save_checkpoint(self, checkpoint_dict: dict[str, Any]) -> None:
        assert self.checkpoint_path is not None, "Checkpoint path is not set but save_checkpoint has been called."
        try:
            # write-then-rename: a pre-emption mid-write must not corrupt the previous round's state
            tmp_path = f"{self.checkpoint_path}.tmp"
            torch.save(checkpoint_dict, tmp_path)
            os.replace(tmp_path, self.checkpoint_path)
        except Exception as exc:
            log(ERROR, f"Encountered the following error while saving the checkpoint: {exc}")
            raise

    def load_checkpoint(self) -> dict[str, Any]:
        assert self.checkpoint_path is not None, "Checkpoint path is not set but load_checkpoint has been called."
        assert self.checkpoint_exists(), f"Could not verify existence of checkpoint file at {self.checkpoint_path}"
        log(INFO, f"Loading state from checkpoint at {self.checkpoint_path}")
        return torch.load(self.checkpoint_path, weights_only=False)

    def add_to_snapshot_attr(self, name: str, snapshotter: AbstractSnapshotter, input_type: Any) -> None:
        self.snapshot_attrs[name] = (snapshotter, input_type)

    def delete_from_snapshot_attr(self, name: str) -> None:
        del self.snapshot_attrs[name]

    def save_state(self) -> None:
        for name, (snapshotter, expected_type) in self.snapshot_attrs.items():
            self.snapshot_ckpt[name] = snapshotter.save_attribute(self._dict_wrap_attr(name, expected_type))
        log(INFO, f"Saving the state to checkpoint at {self.checkpoint_path}")
        self.save_checkpoint(self.snapshot_ckpt)
        self.snapshot_ckpt = {}

    def load_state(self, attributes: list[str] | None = None) -> None:
        assert self.checkpoint_exists(), f"No state checkpoint to load. {self.checkpoint_path} does not exist"
        if attributes is None:
            attributes = list(self.snapshot_attrs.keys())
            if not attributes:
                log(WARNING, "self.snapshot_attrs is empty, which may be undesired behavior.")
        self.snapshot_ckpt = self.load_checkpoint()
        for name in attributes:
            snapshotter, expected_type = self.snapshot_attrs[name]
            wrapped = self._dict_wrap_attr(name, expected_type)
            snapshotter.load_attribute(self.snapshot_ckpt[name], wrapped)
            self.set_attribute(name, wrapped["None"] if list(wrapped.keys()) == ["None"] else wrapped)
        log(INFO, f"Loaded the checkpointed state from {self.checkpoint_path}")
        self.snapshot_ckpt = {}

    @abstractmethod
    def get_attribute(self, name: str) -> Any:
        raise NotImplementedError

    @abstractmethod
    def set_attribute(self, name: str, value: Any) -> None:
        raise NotImplementedError

    def _dict_wrap_attr(self, name: str, expected_type: Any) -> dict[str, Any]:
        attribute = self.get_attribute(name)
        if isinstance(attribute, expected_type):
            return {"None": attribute}
        if isinstance(attribute, dict):
            for key, value in attribute.items():
                if not isinstance(value, expected_type):
                    raise ValueError(f"Incompatible type of attribute {type(attribute)} for key {key}")
            return attribute
        raise ValueError(f"Incompatible type of attribute {type(attribute)}, expected {expected_type}")


def default_client_snapshot_attrs() -> SnapshotSpec:
    return {
        "model": (TorchModuleSnapshotter(), nn.Module),
        "optimizers": (OptimizerSnapshotter(), Optimizer),
        "lr_schedulers": (LRSchedulerSnapshotter(), LRScheduler),
        "total_steps": (SingletonSnapshotter(), int),
        "total_epochs": (SingletonSnapshotter(), int),
        "reports_manager": (SerializableObjectSnapshotter(), ReportsManager),
        "train_loss_meter": (SerializableObjectSnapshotter(), LossMeter),
        "train_metric_manager": (SerializableObjectSnapshotter(), MetricManager),
    }


class ClientStateCheckpointer(StateCheckpointer):
    def __init__(
        self, checkpoint_dir: Path, checkpoint_name: str | None = None, snapshot_attrs: SnapshotSpec | None = None
    ) -> None:
        super().__init__(checkpoint_dir, checkpoint_name, snapshot_attrs or default_client_snapshot_attrs())
        self.client: BasicClient | None = None

    def maybe_set_default_checkpoint_name(self) -> None:
        assert self.client is not None, "Attempting to save client state but client is None"
        if self.checkpoint_name is None:
            self.set_checkpoint_path(self.checkpoint_dir, f"client_{self.client.client_name}_state.pt")

    def save_client_state(self, client: BasicClient) -> None:
        self.client = client
        try:
            self.maybe_set_default_checkpoint_name()
            self.save_state()
        finally:
            self.client = None

    def maybe_load_client_state(self, client: BasicClient, attributes: list[str] | None = None) -> bool:
        self.client = client
        try:
            self.maybe_set_default_checkpoint_name()
            if not self.checkpoint_exists():
                log(INFO, f"No state checkpoint found at: {self.checkpoint_path}")
                return False
            self.load_state(attributes)
            log(INFO, f"State checkpoint successfully loaded from: {self.checkpoint_path}")
            return True
        finally:
            self.client = None

    def get_attribute(self, name: str) -> Any:
        assert self.client is not None, "Client is not set."
        return getattr(self.client, name)

    def set_attribute(self, name: str, value: Any) -> None:
        assert self.client is not None, "Client is not set."
        setattr(self.client, name, value)


def default_server_snapshot_attrs() -> SnapshotSpec:
    return {
        "model": (TorchModuleSnapshotter(), nn.Module),
        "current_round": (SingletonSnapshotter(), int),
        "reports_manager": (SerializableObjectSnapshotter(), ReportsManager),
        "server_name": (StringSnapshotter(), str),
        "history": (HistorySnapshotter(), History),
    }


class ServerStateCheckpointer(StateCheckpointer):
    def __init__(
        self, checkpoint_dir: Path, checkpoint_name: str | None = None, snapshot_attrs: SnapshotSpec | None = None
    ) -> None:
        super().__init__(checkpoint_dir, checkpoint_name, snapshot_attrs or default_server_snapshot_attrs())
        self.server: FlServer | None = None
        self.server_model: nn.Module | None = None

    def maybe_set_default_checkpoint_name(self) -> None:
        assert self.server is not None, "Attempting to save server state but server is None"
        if self.checkpoint_name is None:
            self.set_checkpoint_path(self.checkpoint_dir, f"server_{self.server.server_name}_state.pt")

    def save_server_state(self, server: FlServer, model: nn.Module) -> None:
        self.server, self.server_model = server, model
        try:
            self.maybe_set_default_checkpoint_name()
            self.save_state()
        finally:
            self.server, self.server_model = None, None

    def maybe_load_server_state(
        self, server: FlServer, model: nn.Module, attributes: list[str] | None = None
    ) -> nn.Module | None:
        self.server, self.server_model = server, model
        try:
            self.maybe_set_default_checkpoint_name()
            if not self.checkpoint_exists():
                log(INFO, f"No state checkpoint found at: {self.checkpoint_path}")
                return None
            self.load_state(attributes)
            log(INFO, f"State checkpoint successfully loaded from: {self.checkpoint_path}")
            return self.server_model
        finally:
            self.server = None
            self.server_model = None

    def get_attribute(self, name: str) -> Any:
        assert self.server is not None, "Server is not set."
        return self.server_model if name == "model" else getattr(self.server, name)

    def set_attribute(self, name: str, value: Any) -> None:
        assert self.server is not None, "Server is not set."
        if name == "model":
            self.server_model = value
        else:
            setattr(self.server, name, value)


class NnUnetServerStateCheckpointer(ServerStateCheckpointer):
    def __init__(self, checkpoint_dir: Path, checkpoint_name: str | None = None) -> None:
        attrs = default_server_snapshot_attrs()
        attrs.update(
            {
                "nnunet_plans_bytes": (BytesSnapshotter(), bytes),
                "num_segmentation_heads": (SingletonSnapshotter(), int),
                "num_input_channels": (SingletonSnapshotter(), int),
                "global_deep_supervision": (EnumSnapshotter(), bool),
                "nnunet_config": (EnumSnapshotter(), Enum),
            }
        )
        super().__init__(checkpoint_dir, checkpoint_name, attrs)
