"""Checkpointers for DP-wrapped models (parity: ``fl4health/checkpointing/opacus_checkpointer.py:20-174``).

The on-disk format is a *pickled state_dict* (not a pickled module) — DP wrappers carry hooks/closures that do not
pickle.  ``load_checkpoint`` pours the state into a user-supplied architecture, stripping the ``_module.`` prefix a
grad-sample wrapper adds.
"""

from __future__ import annotations

import pickle
from logging import ERROR, INFO
from typing import Any

import torch
from torch import nn

from fl4health_b200.checkpointing.checkpointer import FunctionTorchModuleCheckpointer
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Scalar


class OpacusCheckpointer(FunctionTorchModuleCheckpointer):
    def maybe_checkpoint(self, model: nn.Module, loss: float, metrics: dict[str, Scalar]) -> None:
        score = self.checkpoint_score_function(loss, metrics)
        if not self._should_checkpoint(score):
            log(
                INFO,
                f"Not checkpointing the model: Current {self.checkpoint_score_name} score ({score}) is not "
                f"{self.comparison_str} Best score ({self.best_score})",
            )
            return
        log(INFO, f"Checkpointing the model state: {self.checkpoint_score_name} score ({score}) -> {self.checkpoint_path}")
        self.best_score = score
        self._save_model_state(model)

    def _save_model_state(self, model: nn.Module) -> None:
        state = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        try:
            with open(self.checkpoint_path, "wb") as handle:
                pickle.dump(state, handle)
        except Exception as exc:
            log(ERROR, f"Encountered the following error while saving the checkpoint: {exc}")
            raise

    def load_checkpoint(self, path_to_checkpoint: str | None = None) -> nn.Module:
        raise NotImplementedError(
            "When loading from Opacus checkpointers, you need to provide a model into which state is loaded. "
            "Please use load_best_checkpoint_into_model instead"
        )

    def load_best_checkpoint(self, model: nn.Module, target_is_grad_sample_module: bool = False) -> None:
        with open(self.checkpoint_path, "rb") as handle:
            state: dict[str, Any] = pickle.load(handle)
        wrapped_keys = all(k.startswith("_module.") for k in state)
        if wrapped_keys and not target_is_grad_sample_module:
            state = {k[len("_module.") :]: v for k, v in state.items()}
        elif not wrapped_keys and target_is_grad_sample_module:
            state = {f"_module.{k}": v for k, v in state.items()}
        model.load_state_dict({k: torch.as_tensor(v) for k, v in state.items()}, strict=True)


    def load_best_checkpoint_into_model(self, target_model: nn.Module, target_is_grad_sample_module: bool = False) -> None:
        """The reference's name for ``load_best_checkpoint`` (``opacus_checkpointer.py:89-107``)."""
        self.load_best_checkpoint(target_model, target_is_grad_sample_module)


class LatestOpacusCheckpointer(OpacusCheckpointer):
    def __init__(self, checkpoint_dir: str, checkpoint_name: str) -> None:
        super().__init__(checkpoint_dir, checkpoint_name, lambda loss, metrics: 0.0, "Latest", False)

    def maybe_checkpoint(self, model: nn.Module, loss: float, metrics: dict[str, Scalar]) -> None:
        log(INFO, f"Saving latest checkpoint with LatestOpacusCheckpointer as {self.checkpoint_path}")
        self._save_model_state(model)


class BestLossOpacusCheckpointer(OpacusCheckpointer):
    def __init__(self, checkpoint_dir: str, checkpoint_name: str) -> None:
        super().__init__(checkpoint_dir, checkpoint_name, lambda loss, metrics: loss, "Loss", False)
