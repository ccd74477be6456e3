"""Encoder-based dimensionality reduction transforms (parity:
``fl4health/preprocessing/autoencoders/dim_reduction.py:9-168``): callables that load a trained (V/CV)AE checkpoint
and map samples (or batches) to latent codes; usable as dataset (batch) transforms."""

from __future__ import annotations

from pathlib import Path

import torch

DEVICE: torch.device = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")


class AutoEncoderProcessing:
    def __init__(self, checkpointing_path: Path, device: torch.device = DEVICE) -> None:
        self.checkpointing_path = checkpointing_path
        self.device = device
        self.load_autoencoder()

    def load_autoencoder(self) -> None:
        autoencoder = torch.load(self.checkpointing_path, weights_only=False)
        autoencoder.eval()
        self.autoencoder = autoencoder.to(self.device)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}"

    def _latent(self, mu: torch.Tensor, logvar: torch.Tensor, mu_only: bool) -> torch.Tensor:
        # cat on the last axis: the latent axis for single samples and for batches alike
        return mu.detach().clone() if mu_only else torch.cat((mu.detach(), logvar.detach()), dim=-1)


class AeProcessor(AutoEncoderProcessing):
    @torch.no_grad()
    def __call__(self, sample: torch.Tensor) -> torch.Tensor:
        return self.autoencoder.encode(sample.to(self.device)).detach().clone()


class VaeProcessor(AutoEncoderProcessing):
    def __init__(self, checkpointing_path: Path, device: torch.device = DEVICE, return_mu_only: bool = False) -> None:
        super().__init__(checkpointing_path, device)
        self.return_mu_only = return_mu_only

    @torch.no_grad()
    def __call__(self, sample: torch.Tensor) -> torch.Tensor:
        mu, logvar = self.autoencoder.encode(sample.to(self.device))
        return self._latent(mu, logvar, self.return_mu_only)


class CvaeFixedConditionProcessor(AutoEncoderProcessing):
    def __init__(self, checkpointing_path: Path, condition: torch.Tensor, device: torch.device = DEVICE,
                 return_mu_only: bool = False) -> None:
        super().__init__(checkpointing_path, device)
        assert condition.dim() == 1, f"Error: condition should be a 1D vector instead of a {condition.dim()}D tensor."
        self.condition = condition
        self.return_mu_only = return_mu_only

    @torch.no_grad()
    def __call__(self, sample: torch.Tensor) -> torch.Tensor:
        condition = self.condition if sample.dim() == 1 else self.condition.expand(sample.shape[0], -1)
        mu, logvar = self.autoencoder.encode(sample.to(self.device), condition.to(self.device))
        return self._latent(mu, logvar, self.return_mu_only)


class CvaeVariableConditionProcessor(AutoEncoderProcessing):
    def __init__(self, checkpointing_path: Path, device: torch.device = DEVICE, return_mu_only: bool = False) -> None:
        super().__init__(checkpointing_path, device)
        self.return_mu_only = return_mu_only

    @torch.no_grad()
    def __call__(self, sample: torch.Tensor, condition: torch.Tensor) -> torch.Tensor:
        if condition.size(0) > 1:
            assert condition.size(0) == sample.size(0), (
                f"Error: Condition shape: {condition.shape} does not match the data shape: {sample.shape}")
        mu, logvar = self.autoencoder.encode(sample.to(self.device), condition.to(self.device))
        return self._latent(mu, logvar, self.return_mu_only)
