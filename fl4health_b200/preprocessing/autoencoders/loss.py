"""VAE objective on the packed model output (parity: ``fl4health/preprocessing/autoencoders/loss.py:8-80``).

The variational models return ``[logvar | mu | flattened reconstruction]`` (see ``model_bases/autoencoders_base.py``);
the loss is ``base_loss(recon, target) + KL(N(mu, exp(logvar)) || N(0, I))`` with the KL summed over batch and latent."""

from __future__ import annotations

import torch
from torch.nn.modules.loss import _Loss

REQUIRED_PREDS_DIMENSIONS = 2


class VaeLoss(_Loss):
    def __init__(self, latent_dim: int, base_loss: _Loss) -> None:
        super().__init__()
        self.base_loss = base_loss
        self.latent_dim = latent_dim

    def standard_normal_kl_divergence_loss(self, mu: torch.Tensor, logvar: torch.Tensor) -> torch.Tensor:
        return -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp())

    def unpack_model_output(self, preds: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        assert preds.dim() == REQUIRED_PREDS_DIMENSIONS, (
            f"Expected a 2D tensor for VaeLoss, but got {preds.dim()}D tensor with shape {preds.shape}.")
        logvar, mu, recon = torch.split(preds, [self.latent_dim, self.latent_dim, preds.shape[1] - 2 * self.latent_dim], dim=1)
        return recon, mu, logvar

    def forward(self, preds: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        recon, mu, logvar = self.unpack_model_output(preds)
        return self.base_loss(recon.reshape(target.shape), target) + self.standard_normal_kl_divergence_loss(mu, logvar)
