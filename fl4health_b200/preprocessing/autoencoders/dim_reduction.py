"""Encoder-based dimensionality reduction transforms (parity:
``fl4health/preprocessing/autoencoders/dim_reduction.py:9-168``): callables that load a trained (V/CV)AE checkpoint
and map samples (or batches) to latent codes; usable as dataset (batch) transforms.

All variational processors are one mechanism — encode with zero or more conditioning tensors, return ``mu`` or
``[mu | logvar]`` — and differ only in where the condition comes from (none / fixed at construction / per call)."""

from __future__ import annotations

from pathlib import Path

import torch

DEVICE: torch.device = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")


class AutoEncoderProcessing:
    def __init__(self, checkpointing_path: Path, device: torch.device = DEVICE) -> None:
        self.checkpointing_path, self.device = checkpointing_path, device
        self.load_autoencoder()

    def load_autoencoder(self) -> None:
        self.autoencoder = torch.load(self.checkpointing_path, weights_only=False).eval().to(self.device)

    def __repr__(self) -> str:
        return type(self).__name__

    @torch.no_grad()
    def _latent_code(self, sample: torch.Tensor, *conditions: torch.Tensor, mu_only: bool) -> torch.Tensor:
        mu, logvar = self.autoencoder.encode(sample.to(self.device), *(c.to(self.device) for c in conditions))
        # the latent axis is the last one for single samples and for batches alike
        return mu.detach().clone() if mu_only else torch.cat((mu.detach(), logvar.detach()), dim=-1)


class AeProcessor(AutoEncoderProcessing):
    @torch.no_grad()
    def __call__(self, sample: torch.Tensor) -> torch.Tensor:
        return self.autoencoder.encode(sample.to(self.device)).detach().clone()


class VaeProcessor(AutoEncoderProcessing):
    def __init__(self, checkpointing_path: Path, device: torch.device = DEVICE, return_mu_only: bool = False) -> None:
        super().__init__(checkpointing_path, device)
        self.return_mu_only = return_mu_only

    def __call__(self, sample: torch.Tensor) -> torch.Tensor:
        return self._latent_code(sample, mu_only=self.return_mu_only)


class CvaeFixedConditionProcessor(AutoEncoderProcessing):
    def __init__(self, checkpointing_path: Path, condition: torch.Tensor, device: torch.device = DEVICE,
                 return_mu_only: bool = False) -> None:
        super().__init__(checkpointing_path, device)
        assert condition.dim() == 1, f"Error: condition should be a 1D vector instead of a {condition.dim()}D tensor."
        self.condition, self.return_mu_only = condition, return_mu_only

    def __call__(self, sample: torch.Tensor) -> torch.Tensor:
        batched = sample.dim() > 1  # one condition vector serves every row of a batch
        condition = self.condition.expand(sample.shape[0], -1) if batched else self.condition
        return self._latent_code(sample, condition, mu_only=self.return_mu_only)


class CvaeVariableConditionProcessor(AutoEncoderProcessing):
    def __init__(self, checkpointing_path: Path, device: torch.device = DEVICE, return_mu_only: bool = False) -> None:
        super().__init__(checkpointing_path, device)
        self.return_mu_only = return_mu_only

    def __call__(self, sample: torch.Tensor, condition: torch.Tensor) -> torch.Tensor:
        if condition.size(0) > 1 and condition.size(0) != sample.size(0):
            raise AssertionError(f"Error: Condition shape: {condition.shape} does not match the data shape: {sample.shape}")
        return self._latent_code(sample, condition, mu_only=self.return_mu_only)
