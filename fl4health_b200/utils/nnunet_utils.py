"""nnU-Net helpers that do not need ``nnunetv2`` itself (parity: ``fl4health/utils/nnunet_utils.py:40-592``):
config enum, deep-supervision list<->dict conversion, segmentation post-processing, poly LR schedulers, an iterator
wrapper for nnU-Net's infinite batch generators, stdout-to-logger stream."""

from __future__ import annotations

import io
import os
import signal
from collections.abc import Callable, Iterator, Sequence
from enum import Enum
from logging import INFO, WARNING, Logger
from math import ceil
from typing import Any

import numpy as np
import torch
from torch import nn
from torch.nn.modules.loss import _Loss
from torch.optim.lr_scheduler import _LRScheduler

from fl4health_b200.common.logger import log


class NnunetConfig(Enum):
    """Model configurations of nnunetv2 (2.5.x)."""

    _2D = "2d"
    _3D_FULLRES = "3d_fullres"
    _3D_CASCADE = "3d_cascade_fullres"
    _3D_LOWRES = "3d_lowres"


NNUNET_DEFAULT_NP = {NnunetConfig._2D: 8, NnunetConfig._3D_FULLRES: 4, NnunetConfig._3D_CASCADE: 4, NnunetConfig._3D_LOWRES: 8}
NNUNET_N_SPATIAL_DIMS = {NnunetConfig._2D: 2, NnunetConfig._3D_FULLRES: 3, NnunetConfig._3D_CASCADE: 3, NnunetConfig._3D_LOWRES: 3}


def use_default_signal_handlers(fn: Callable) -> Callable:
    """Run ``fn`` with the default SIGINT/SIGTERM handlers (nnU-Net's multiprocess augmenters install their own
    children; a server-installed handler inherited by them would swallow shutdown signals), then restore."""

    def new_fn(*args: Any, **kwargs: Any) -> Any:
        try:
            previous = (signal.getsignal(signal.SIGINT), signal.getsignal(signal.SIGTERM))
            signal.signal(signal.SIGINT, signal.default_int_handler)
            signal.signal(signal.SIGTERM, signal.SIG_DFL)
        except ValueError:  # not on the main thread: nothing to swap
            return fn(*args, **kwargs)
        try:
            return fn(*args, **kwargs)
        finally:
            signal.signal(signal.SIGINT, previous[0])
            signal.signal(signal.SIGTERM, previous[1])

    return new_fn


def set_nnunet_env(verbose: bool = False, **kwargs: Any) -> None:
    """Set nnU-Net environment variables (``nnUNet_raw``, ``nnUNet_preprocessed``, ``nnUNet_results``, ...)."""
    for key, value in kwargs.items():
        os.environ[key] = str(value)
        if verbose:
            log(INFO, f"Resetting env var '{key}' to '{value}'")


def reload_modules(packages: Sequence[str]) -> None:
    """Reload already imported modules whose dotted name starts with one of ``packages`` (nnU-Net reads its env vars
    at import time)."""
    import importlib
    import sys

    for name in [m for m in list(sys.modules) if any(m == p or m.startswith(p + ".") for p in packages)]:
        try:
            importlib.reload(sys.modules[name])
        except Exception as exc:  # noqa: BLE001
            log(WARNING, f"Could not reload {name}: {exc}")


def set_nnunet_env_and_reload_modules(verbose: bool = False, **kwargs: Any) -> None:
    set_nnunet_env(verbose, **kwargs)
    reload_modules(["nnunetv2", "fl4health_b200.clients.nnunet_client"])


def convert_deep_supervision_list_to_dict(tensor_list: list[torch.Tensor] | tuple[torch.Tensor, ...], num_spatial_dims: int) -> dict[str, torch.Tensor]:
    """Keys are ``"{index}-{d0}x{d1}[x{d2}]"``: index keeps nnU-Net's resolution order, the suffix documents the scale."""
    return {f"{i}-" + "x".join(str(s) for s in t.shape[-num_spatial_dims:]): t for i, t in enumerate(tensor_list)}


def convert_deep_supervision_dict_to_list(tensor_dict: dict[str, torch.Tensor]) -> list[torch.Tensor]:
    return [t for _, t in sorted(tensor_dict.items(), key=lambda kv: int(kv[0].split("-")[0]))]


def get_segs_from_probs(preds: torch.Tensor, has_regions: bool = False, threshold: float = 0.5) -> torch.Tensor:
    """Hard one-hot segmentation from (soft) predictions ``[B, C, ...]``; with regions, classes may overlap and are
    thresholded individually, masked by "not background" (channel 0)."""
    if has_regions:
        segs = preds > threshold
        return segs * ~segs[:, 0]
    winners = preds.argmax(1, keepdim=True)
    return torch.zeros(preds.shape, device=preds.device, dtype=torch.float32).scatter_(1, winners, 1).long()


def collapse_one_hot_tensor(input: torch.Tensor, dim: int = 0) -> torch.Tensor:  # noqa: A002
    return torch.argmax(input.long(), dim=dim).to(input.device)


def get_dataset_n_voxels(source_plans: dict, n_cases: int) -> float:
    configs = source_plans["configurations"]
    cfg = configs[NnunetConfig._3D_FULLRES.value] if NnunetConfig._3D_FULLRES.value in configs else configs[NnunetConfig._2D.value]
    return float(np.prod(cfg["median_image_size_in_voxels"], dtype=np.float64) * n_cases)


def prepare_loss_arg(tensor: torch.Tensor | dict[str, torch.Tensor]) -> torch.Tensor | list[torch.Tensor]:
    """Tensor stays; a dict with several entries is deep supervision (-> list in resolution order); a singleton dict is
    unwrapped."""
    if isinstance(tensor, torch.Tensor):
        return tensor
    if isinstance(tensor, dict):
        return convert_deep_supervision_dict_to_list(tensor) if len(tensor) > 1 else next(iter(tensor.values()))
    raise ValueError(f"Unrecognized type for tensor: {type(tensor)}")


class NnUNetDataLoaderWrapper:
    """Finite-epoch view over nnU-Net's infinite augmenter (``{"data": ..., "target": ...}`` batches).

    ``len`` = ``ceil(dataset voxels / voxels per batch)`` like the reference (one "epoch" sees roughly every voxel
    once); with deep supervision the list of targets becomes the keyed dict the client expects."""

    def __init__(self, nnunet_augmenter: Any, nnunet_config: NnunetConfig | str, infinite: bool = False,
                 set_len: int | None = None, ref_image_shape: Sequence[int] | None = None, n_cases: int | None = None) -> None:
        self.nnunet_augmenter = nnunet_augmenter
        self.nnunet_config = NnunetConfig(nnunet_config) if isinstance(nnunet_config, str) else nnunet_config
        self.num_spatial_dims = NNUNET_N_SPATIAL_DIMS[self.nnunet_config]
        self.infinite = infinite
        self.set_len = set_len
        self.ref_image_shape, self.n_cases = ref_image_shape, n_cases
        self.current_step = 0
        self._iterator: Iterator | None = None
        generator = getattr(nnunet_augmenter, "generator", nnunet_augmenter)
        self._source_dataset = getattr(generator, "_data", None)

    @property
    def dataset(self) -> Any:
        """What ``len(loader.dataset)`` reports as the client's sample count: nnU-Net's case list when available,
        otherwise the number of samples one finite pass yields."""
        if self._source_dataset is not None:
            return self._source_dataset
        generator = getattr(self.nnunet_augmenter, "generator", self.nnunet_augmenter)
        return range(len(self) * int(getattr(generator, "batch_size", 1)))

    def __next__(self) -> tuple[torch.Tensor, torch.Tensor | dict[str, torch.Tensor]]:
        if not self.infinite and self.current_step == len(self):
            self.reset()
            raise StopIteration
        self.current_step += 1
        if self._iterator is None:
            self._iterator = iter(self.nnunet_augmenter)
        batch = next(self._iterator)
        data, target = batch["data"], batch["target"]
        if isinstance(target, (list, tuple)):
            return data, convert_deep_supervision_list_to_dict(list(target), self.num_spatial_dims)
        if isinstance(target, torch.Tensor):
            return data, target
        raise TypeError("Was expecting nnunet target to be a tensor or a list/tuple of tensors")

    def __len__(self) -> int:
        if self.set_len is not None:
            return self.set_len
        generator = getattr(self.nnunet_augmenter, "generator", self.nnunet_augmenter)
        patch = getattr(generator, "final_patch_size", None)
        batch_size = getattr(generator, "batch_size", 1)
        if patch is not None and self.ref_image_shape is not None and self.n_cases is not None:
            voxels = float(np.prod(self.ref_image_shape, dtype=np.float64) * self.n_cases)
            return max(1, ceil(voxels / float(np.prod(patch) * batch_size)))
        if self._source_dataset is not None and hasattr(self._source_dataset, "__len__"):
            return max(1, ceil(len(self._source_dataset) / batch_size))
        raise ValueError("Cannot infer an epoch length: pass set_len or (ref_image_shape, n_cases)")

    def reset(self) -> None:
        self.current_step = 0

    def __iter__(self) -> NnUNetDataLoaderWrapper:
        self.reset()
        return self

    def shutdown(self) -> None:
        finish = getattr(self.nnunet_augmenter, "_finish", None)
        if callable(finish):
            finish()


class Module2LossWrapper(_Loss):
    """Adapts an ``nn.Module`` loss to the ``_Loss`` type the clients annotate."""

    def __init__(self, loss: nn.Module, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self.loss = loss

    def forward(self, pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return self.loss(pred, target)


class StreamToLogger(io.StringIO):
    """File-like object that forwards complete lines to a logger (used to demote nnU-Net's prints to DEBUG)."""

    def __init__(self, logger: Logger, level: int) -> None:
        super().__init__()
        self.logger, self.level = logger, level
        self.linebuf = ""

    def write(self, buf: str) -> int:
        for line in buf.rstrip().splitlines():
            self.logger.log(self.level, line.rstrip())
        return len(buf)

    def flush(self) -> None:
        pass


class PolyLRSchedulerWrapper(_LRScheduler):
    """Polynomial decay held constant over windows of ``steps_per_lr`` steps (nnU-Net decays per epoch of 250 steps):
    ``lr = lr0 (1 - window / n_windows)^exponent``."""

    def __init__(self, optimizer: torch.optim.Optimizer, initial_lr: float, max_steps: int, exponent: float = 0.9,
                 steps_per_lr: int = 250) -> None:
        self.initial_lr, self.max_steps, self.exponent, self.steps_per_lr = initial_lr, max_steps, exponent, steps_per_lr
        self.num_windows = ceil(max_steps / steps_per_lr)
        self._step_count: int
        super().__init__(optimizer, -1)

    def get_lr(self) -> list[float]:  # type: ignore[override]
        if self._step_count - 1 == self.max_steps + 1:
            log(WARNING, f"Current LR step of {self._step_count} reached Max Steps of {self.max_steps}. LR will remain fixed.")
        step = min(self._step_count - 1, self.max_steps)
        window = int(step / self.steps_per_lr)
        new_lr = self.initial_lr * (1 - window / self.num_windows) ** self.exponent
        if step % self.steps_per_lr == 0 and step not in {0, self.max_steps}:
            log(INFO, f"Decaying LR of optimizer to {new_lr} at step {step}")
        return [new_lr] * len(self.optimizer.param_groups)


class LocalPolyLRScheduler(_LRScheduler):
    """nnU-Net's own ``PolyLRScheduler``: ``lr = lr0 (1 - step / max_steps)^exponent`` on every ``step()``."""

    def __init__(self, optimizer: torch.optim.Optimizer, initial_lr: float, max_steps: int, exponent: float = 0.9,
                 current_step: int | None = None) -> None:
        self.initial_lr, self.max_steps, self.exponent = initial_lr, max_steps, exponent
        self.ctr = 0
        super().__init__(optimizer, current_step if current_step is not None else -1)

    def step(self, current_step: int | None = None) -> None:  # type: ignore[override]
        if current_step is None or current_step == -1:
            current_step = self.ctr
            self.ctr += 1
        new_lr = self.initial_lr * (1 - current_step / self.max_steps) ** self.exponent
        for group in self.optimizer.param_groups:
            group["lr"] = new_lr
