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
        self.latent_dim, self.base_loss = latent_dim, base_loss

    def unpack_model_output(self, preds: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """(reconstruction, mu, logvar) views of the packed prediction."""
        if preds.dim() != REQUIRED_PREDS_DIMENSIONS:
            raise AssertionError(f"Expected a 2D tensor for VaeLoss, but got {preds.dim()}D tensor with shape {preds.shape}.")
        d = self.latent_dim
        return preds[:, 2 * d:], preds[:, d : 2 * d], preds[:, :d]

    def standard_normal_kl_divergence_loss(self, mu: torch.Tensor, logvar: torch.Tensor) -> torch.Tensor:
        return 0.5 * (mu.square() + logvar.exp() - logvar - 1).sum()

    def forward(self, preds: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        reconstruction, mu, logvar = self.unpack_model_output(preds)
        fidelity = self.base_loss(reconstruction.reshape(target.shape), target)
        return fidelity + self.standard_normal_kl_divergence_loss(mu, logvar)
