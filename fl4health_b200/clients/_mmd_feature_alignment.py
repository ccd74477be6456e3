"""Shared machinery of the feature-space MMD clients (Ditto / MR-MTL x MK-MMD / Deep-MMD).

The four reference clients (``fl4health/clients/mkmmd_clients/*.py``, ``deep_mmd_clients/*.py``) each re-implement
the same ~250 lines: forward hooks on the personal model and on a frozen copy of the round-start global model,
a per-layer MMD penalty between the two feature batches, and a periodic refresh of the kernel (MK-MMD: the beta
weights via a QP; Deep-MMD: a few AdamW steps of the learned kernel) on features accumulated over the training set.
Here that logic lives once, in ``MmdFeatureAlignmentMixin``; the concrete clients only choose the base client, the
loss family and where the frozen anchor model comes from.
"""

from __future__ import annotations

import dataclasses
from logging import ERROR, INFO

import torch
from torch import nn

from fl4health_b200.checkpointing.client_module import CheckpointMode
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Scalar
from fl4health_b200.model_bases.feature_extractor_buffer import FeatureExtractorBuffer
from fl4health_b200.utils.typing import TorchFeatureType, TorchInputType

INIT_GLOBAL_PREFIX = "init_global"


class MmdFeatureAlignmentMixin:
    """Expects the host class to be a ``BasicClient`` descendant that owns ``self.model`` and provides a frozen
    ``self.initial_global_model`` by the time ``_attach_anchor_hooks`` is called (each round)."""

    mmd_loss_name: str  # "mkmmd_loss" | "deep_mmd_loss": prefix of the reported additional losses

    def _init_mmd(self, weight: float, layers: dict[str, bool], losses: dict[str, nn.Module], refresh_interval: int,
                  num_accumulating_batches: int | None, what: str) -> None:
        self.mmd_weight = weight
        if weight == 0:
            log(ERROR, f"{what} loss weight is set to 0. As the {what} loss will not be computed, use the vanilla client instead.")
        if refresh_interval < -1:
            raise ValueError("Invalid kernel refresh interval. It should be either -1, 0 or a positive integer.")
        self.flatten_feature_extraction_layers = layers
        self.mmd_losses = losses
        self.mmd_refresh_interval = refresh_interval
        self.num_accumulating_batches = num_accumulating_batches
        self.local_feature_extractor: FeatureExtractorBuffer
        self.initial_global_feature_extractor: FeatureExtractorBuffer
        # hooks + host-side kernel refreshes re-bind tensors between steps: run the step eagerly
        self.engine = dataclasses.replace(self.engine, cuda_graphs=False)  # type: ignore[has-type]

    # ---------------------------------------------------------------------------------------------- hooks
    def _attach_local_hooks(self) -> None:
        self.local_feature_extractor = FeatureExtractorBuffer(self.model, self.flatten_feature_extraction_layers)  # type: ignore[attr-defined]
        self.local_feature_extractor._maybe_register_hooks()

    def _attach_anchor_hooks(self) -> None:
        self.initial_global_feature_extractor = FeatureExtractorBuffer(
            self.initial_global_model, self.flatten_feature_extraction_layers  # type: ignore[attr-defined]
        )
        self.initial_global_feature_extractor._maybe_register_hooks()

    def _maybe_checkpoint(self, loss: float, metrics: dict[str, Scalar], checkpoint_mode: CheckpointMode) -> None:
        """Hook closures are not picklable: detach them around checkpointing."""
        self.local_feature_extractor.remove_hooks()
        super()._maybe_checkpoint(loss=loss, metrics=metrics, checkpoint_mode=checkpoint_mode)  # type: ignore[misc]
        self.local_feature_extractor._maybe_register_hooks()

    # ------------------------------------------------------------------------------------ kernel refresh
    def _should_optimize_betas(self, step: int) -> bool:
        return (
            (step - 1) % self.mmd_refresh_interval == 0
            and getattr(self, "initial_global_model", None) is not None
            and self.mmd_weight != 0
        )

    def _refresh_kernel(self, layer: str, local: torch.Tensor, anchor: torch.Tensor) -> None:
        raise NotImplementedError

    def update_after_step(self, step: int, current_round: int | None = None) -> None:
        if self.mmd_refresh_interval > 0 and self._should_optimize_betas(step):
            local, anchor = self.update_buffers(self.model, self.initial_global_model)  # type: ignore[attr-defined]
            for layer in self.mmd_losses:
                self._refresh_kernel(layer, local[layer], anchor[layer])
        super().update_after_step(step, current_round)  # type: ignore[misc]

    def update_buffers(
        self, local_model: nn.Module, initial_global_model: nn.Module
    ) -> tuple[dict[str, torch.Tensor], dict[str, torch.Tensor]]:
        """Features of both models over (a prefix of) the training set, computed in eval mode without autograd."""
        extractors = (self.local_feature_extractor, self.initial_global_feature_extractor)
        for extractor in extractors:
            extractor.clear_buffers()
            extractor.enable_accumulating_features()
        was_training = local_model.training
        local_model.eval()
        assert not initial_global_model.training
        with torch.no_grad():
            for i, (input, _) in enumerate(self.train_loader):  # type: ignore[attr-defined]
                input = input.to(self.device) if isinstance(input, torch.Tensor) else {k: v.to(self.device) for k, v in input.items()}  # type: ignore[attr-defined]
                if isinstance(input, dict):
                    local_model(**input)
                    initial_global_model(**input)
                else:
                    local_model(input)
                    initial_global_model(input)
                if i == self.num_accumulating_batches:
                    break
        local = self.local_feature_extractor.get_extracted_features()
        anchor = self.initial_global_feature_extractor.get_extracted_features()
        if was_training:
            local_model.train()
        for extractor in extractors:
            extractor.disable_accumulating_features()
            extractor.clear_buffers()
        return local, anchor

    # ----------------------------------------------------------------------------------------- step math
    def _collect_features(self, input: TorchInputType) -> TorchFeatureType:
        """Call right after the personal model's forward: its hooked features + (if weighted) the anchor's."""
        features = self.local_feature_extractor.get_extracted_features()
        if self.mmd_weight != 0:
            with torch.no_grad():
                if isinstance(input, dict):
                    self.initial_global_model(**input)  # type: ignore[attr-defined]
                else:
                    self.initial_global_model(input)  # type: ignore[attr-defined]
            for key, value in self.initial_global_feature_extractor.get_extracted_features().items():
                features[f"{INIT_GLOBAL_PREFIX} {key}"] = value
        return features

    def _per_batch_refresh(self, features: TorchFeatureType) -> None:
        """Interval -1: refresh on every batch from the batch's own features (MK-MMD only)."""

    def _mmd_terms(self, features: TorchFeatureType) -> dict[str, torch.Tensor]:
        if self.mmd_weight == 0:
            return {}
        if self.mmd_refresh_interval == -1:
            self._per_batch_refresh(features)
        terms: dict[str, torch.Tensor] = {}
        total = torch.zeros((), device=self.device)  # type: ignore[attr-defined]
        for layer, loss_fn in self.mmd_losses.items():
            value = loss_fn(features[layer], features[f"{INIT_GLOBAL_PREFIX} {layer}"])
            terms[f"{self.mmd_loss_name}_{layer}"] = value.clone()
            total = total + value
        terms[f"{self.mmd_loss_name}_total"] = self.mmd_weight * total
        return terms


class MkMmdMixin(MmdFeatureAlignmentMixin):
    mmd_loss_name = "mkmmd_loss"

    def _init_mkmmd(self, mkmmd_loss_weight: float, feature_extraction_layers, feature_l2_norm_weight: float,  # noqa: ANN001
                    beta_global_update_interval: int, num_accumulating_batches: int | None) -> None:
        from fl4health_b200.losses.mkmmd_loss import MkMmdLoss

        layers = dict.fromkeys(feature_extraction_layers, True) if feature_extraction_layers else {}
        losses = {
            layer: MkMmdLoss(device=self.device, minimize_type_two_error=True, normalize_features=True, layer_name=layer)  # type: ignore[attr-defined]
            for layer in layers
        }
        if beta_global_update_interval == -1:
            log(INFO, "Betas for the MK-MMD loss will be updated for each individual batch.")
        elif beta_global_update_interval == 0:
            log(INFO, "Betas for the MK-MMD loss will not be updated.")
        elif beta_global_update_interval > 0:
            log(INFO, f"Betas for the MK-MMD loss will be updated every {beta_global_update_interval} steps.")
        self._init_mmd(mkmmd_loss_weight, layers, losses, beta_global_update_interval, num_accumulating_batches, "MK-MMD")
        self.mkmmd_loss_weight = mkmmd_loss_weight
        self.mkmmd_losses = losses
        self.feature_l2_norm_weight = feature_l2_norm_weight
        self.beta_global_update_interval = beta_global_update_interval

    def _refresh_kernel(self, layer: str, local: torch.Tensor, anchor: torch.Tensor) -> None:
        loss = self.mkmmd_losses[layer]
        loss.betas = loss.optimize_betas(x=local, y=anchor, lambda_m=1e-5)

    def _per_batch_refresh(self, features: TorchFeatureType) -> None:
        for layer in self.mkmmd_losses:
            self._refresh_kernel(layer, features[layer].detach(), features[f"{INIT_GLOBAL_PREFIX} {layer}"].detach())

    def _feature_norm_term(self, features: TorchFeatureType) -> dict[str, torch.Tensor]:
        if self.feature_l2_norm_weight == 0:
            return {}
        feats = features["features"]
        return {"feature_l2_norm_loss": self.feature_l2_norm_weight * torch.linalg.norm(feats) / len(feats)}


class DeepMmdMixin(MmdFeatureAlignmentMixin):
    mmd_loss_name = "deep_mmd_loss"

    def _init_deep_mmd(self, deep_mmd_loss_weight: float, feature_extraction_layers_with_size: dict[str, int] | None,
                       mmd_kernel_train_interval: int, num_accumulating_batches: int | None) -> None:
        from fl4health_b200.losses.deep_mmd_loss import DeepMmdLoss
        from fl4health_b200.utils.random import restore_random_state, save_random_state

        sizes = feature_extraction_layers_with_size or {}
        layers = dict.fromkeys(sizes.keys(), True)
        state = save_random_state()  # kernel initialisation must not perturb the client's RNG stream
        losses = {layer: DeepMmdLoss(device=self.device, input_size=size) for layer, size in sizes.items()}  # type: ignore[attr-defined]
        restore_random_state(*state)
        self._init_mmd(deep_mmd_loss_weight, layers, losses, mmd_kernel_train_interval, num_accumulating_batches, "Deep MMD")
        self.deep_mmd_loss_weight = deep_mmd_loss_weight
        self.deep_mmd_losses = losses
        self.mmd_kernel_train_interval = mmd_kernel_train_interval

    def _set_kernel_training(self, flag: bool) -> None:
        for loss in self.deep_mmd_losses.values():
            loss.training = flag

    def _refresh_kernel(self, layer: str, local: torch.Tensor, anchor: torch.Tensor) -> None:
        loss = self.deep_mmd_losses[layer]
        loss.training = True
        loss(local, anchor)  # trains the deep kernel for `optimization_steps`
        loss.training = False

    def _per_batch_refresh(self, features: TorchFeatureType) -> None:
        self._set_kernel_training(True)  # interval -1: the loss call itself trains the kernel on this batch
