"""Model-artifact checkpointers.

Parity: ``fl4health/checkpointing/checkpointer.py:15-311`` — same class names, constructor arguments, file naming
(``os.path.join(checkpoint_dir, checkpoint_name)``) and on-disk format (a whole pickled ``nn.Module`` written with
``torch.save`` and read with ``torch.load(weights_only=False)``), so artifacts are interchangeable.

Arena-backed models (parameters are views into one flat device buffer) are materialised into an ordinary standalone
CPU module before pickling so that a checkpoint loads in vanilla PyTorch with per-parameter storages.

Quirk handling: the reference tests ``if self.best_score:`` which treats a best score of exactly 0.0 as "unset"
(:115).  That is fixed here (``is not None``); set ``FL4H_COMPAT_FALSY_BEST_SCORE=1`` to restore the quirk.
"""

from __future__ import annotations

import copy
import os
from abc import ABC, abstractmethod
from collections.abc import Callable
from logging import ERROR, INFO, WARNING

import torch
from torch import nn

from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Scalar

CheckpointScoreFunctionType = Callable[[float, dict[str, Scalar]], float]


def materialize_module(model: nn.Module) -> nn.Module:
    """Deep copy whose tensors own their (CPU) storage — safe to pickle regardless of how ``model`` is stored."""
    clone = copy.deepcopy(model)
    from fl4health_b200.parallel.arena import arena_of

    arena = arena_of(model)
    masters = arena.shadow_names if arena is not None else set()
    with torch.no_grad():
        for name, param in clone.named_parameters():
            if name in masters:  # low-precision compute view: checkpoint the fp32 master instead
                param.data = arena.view(name).detach().cpu().clone().contiguous()
            else:
                param.data = param.data.detach().cpu().clone()
            param.grad = None
        for module in clone.modules():
            for name, buf in list(module._buffers.items()):
                if buf is not None:
                    module._buffers[name] = buf.detach().cpu().clone()
    return clone


def save_module(model: nn.Module, path: str) -> None:
    try:
        torch.save(materialize_module(model), path)
    except Exception as exc:
        log(ERROR, f"Encountered the following error while saving the checkpoint: {exc}")
        raise


class TorchModuleCheckpointer(ABC):
    def __init__(self, checkpoint_dir: str, checkpoint_name: str) -> None:
        self.checkpoint_path = os.path.join(checkpoint_dir, checkpoint_name)

    @abstractmethod
    def maybe_checkpoint(self, model: nn.Module, loss: float, metrics: dict[str, Scalar]) -> None:
        raise NotImplementedError("maybe_checkpoint must be implemented by inheriting classes")

    def load_checkpoint(self, path_to_checkpoint: str | None = None) -> nn.Module:
        return torch.load(path_to_checkpoint or self.checkpoint_path, weights_only=False)


class FunctionTorchModuleCheckpointer(TorchModuleCheckpointer):
    """Keeps the model whose ``checkpoint_score_function(loss, metrics)`` is best so far."""

    def __init__(
        self,
        checkpoint_dir: str,
        checkpoint_name: str,
        checkpoint_score_function: CheckpointScoreFunctionType,
        checkpoint_score_name: str | None = None,
        maximize: bool = False,
    ) -> None:
        super().__init__(checkpoint_dir, checkpoint_name)
        self.best_score: float | None = None
        self.checkpoint_score_function = checkpoint_score_function
        if checkpoint_score_name is None:
            checkpoint_score_name = getattr(checkpoint_score_function, "__name__", "score")
            log(WARNING, f"No checkpoint_score_name provided. Name will default to {checkpoint_score_name}")
        self.checkpoint_score_name = checkpoint_score_name
        self.maximize = maximize
        self.comparison_str = ">=" if maximize else "<="

    def _should_checkpoint(self, comparison_score: float) -> bool:
        unset = (
            not self.best_score if os.environ.get("FL4H_COMPAT_FALSY_BEST_SCORE") == "1" else self.best_score is None
        )
        if unset:
            return True
        assert self.best_score is not None
        return self.best_score <= comparison_score if self.maximize else self.best_score >= comparison_score

    def maybe_checkpoint(self, model: nn.Module, loss: float, metrics: dict[str, Scalar]) -> None:
        score = self.checkpoint_score_function(loss, metrics)
        if not self._should_checkpoint(score):
            log(
                INFO,
                f"Not checkpointing the model: Current {self.checkpoint_score_name} score ({score}) is not "
                f"{self.comparison_str} Best score ({self.best_score})",
            )
            return
        log(
            INFO,
            f"Checkpointing the model: Current {self.checkpoint_score_name} score ({score}) "
            f"{self.comparison_str} Best score ({self.best_score}); saving as {self.checkpoint_path}",
        )
        self.best_score = score
        save_module(model, self.checkpoint_path)


class LatestTorchModuleCheckpointer(FunctionTorchModuleCheckpointer):
    def __init__(self, checkpoint_dir: str, checkpoint_name: str) -> None:
        super().__init__(checkpoint_dir, checkpoint_name, lambda loss, metrics: 0.0, "Latest", False)

    def maybe_checkpoint(self, model: nn.Module, loss: float, _: dict[str, Scalar]) -> None:
        log(INFO, f"Saving latest checkpoint with LatestTorchCheckpointer as {self.checkpoint_path}")
        save_module(model, self.checkpoint_path)


class BestLossTorchModuleCheckpointer(FunctionTorchModuleCheckpointer):
    def __init__(self, checkpoint_dir: str, checkpoint_name: str) -> None:
        super().__init__(checkpoint_dir, checkpoint_name, lambda loss, metrics: loss, "Loss", False)


class BestMetricTorchModuleCheckpointer(FunctionTorchModuleCheckpointer):
    def __init__(
        self,
        checkpoint_dir: str,
        checkpoint_name: str,
        metric: str,
        prefix: str = "val - prediction - ",
        maximize: bool = False,
    ) -> None:
        self.metric_key = f"{prefix}{metric}"

        def metric_score_function(_: float, metrics: dict[str, Scalar]) -> float:
            if self.metric_key not in metrics:
                log(ERROR, f"Could not find '{self.metric_key}' in metrics dict. Available keys are: {metrics.keys()}")
                raise KeyError(self.metric_key)
            try:
                return float(metrics[self.metric_key])  # type: ignore[arg-type]
            except (ValueError, TypeError):
                log(ERROR, f"Could not convert {self.metric_key} into a float score for best metric checkpointing.")
                raise

        super().__init__(checkpoint_dir, checkpoint_name, metric_score_function, metric, maximize)
