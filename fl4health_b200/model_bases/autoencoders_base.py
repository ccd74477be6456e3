"""Auto-encoder wrappers (parity: ``fl4health/model_bases/autoencoders_base.py:8-282``).

The variational models return ONE packed tensor ``[logvar | mu | flattened reconstruction]`` so they fit the
single-tensor prediction contract of the clients; ``preprocessing/autoencoders/loss.py::VaeLoss`` unpacks it.
Both variational flavours share ``_Variational``: encode (with optional conditioning tensors) -> reparameterised sample
-> decode -> pack.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from collections.abc import Callable

import torch
from torch import nn


class AbstractAe(nn.Module, ABC):
    def __init__(self, encoder: nn.Module, decoder: nn.Module) -> None:
        super().__init__()
        self.encoder, self.decoder = encoder, decoder

    @abstractmethod
    def forward(self, input: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError


class BasicAe(AbstractAe):
    def encode(self, input: torch.Tensor) -> torch.Tensor:
        return self.encoder(input)

    def decode(self, latent_vector: torch.Tensor) -> torch.Tensor:
        return self.decoder(latent_vector)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return self.decoder(self.encoder(input))


class _Variational(AbstractAe):
    """Shared body of the (conditional) variational auto-encoders; ``context`` = the conditioning tensors (possibly none)
    handed to both halves of the network."""

    def sampling(self, mu: torch.Tensor, logvar: torch.Tensor) -> torch.Tensor:
        """z = mu + eps * sigma with eps ~ N(0, I) (one fused addcmul)."""
        return torch.addcmul(mu, torch.randn_like(mu), torch.exp(0.5 * logvar))

    def _packed_forward(self, sample: torch.Tensor, *context: torch.Tensor) -> torch.Tensor:
        mu, logvar = self.encoder(sample, *context)
        reconstruction = self.decoder(self.sampling(mu, logvar), *context)
        return torch.cat((logvar, mu, reconstruction.flatten(start_dim=1)), dim=1)


class VariationalAe(_Variational):
    """``encoder(x) -> (mu, logvar)``; forward returns the packed ``[logvar | mu | recon]`` tensor."""

    def encode(self, input: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        mu, logvar = self.encoder(input)
        return mu, logvar

    def decode(self, latent_vector: torch.Tensor) -> torch.Tensor:
        return self.decoder(latent_vector)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return self._packed_forward(input)


class ConditionalVae(_Variational):
    """CVAE: encoder and decoder both receive the condition; ``unpack_input_condition`` splits the single input tensor
    the data loader provides into (input, condition)."""

    def __init__(
        self, encoder: nn.Module, decoder: nn.Module,
        unpack_input_condition: Callable[[torch.Tensor], tuple[torch.Tensor, torch.Tensor]] | None = None,
    ) -> None:
        super().__init__(encoder, decoder)
        self.unpack_input_condition = unpack_input_condition

    def encode(self, input: torch.Tensor, condition: torch.Tensor | None = None) -> tuple[torch.Tensor, torch.Tensor]:
        mu, logvar = self.encoder(input, condition)
        return mu, logvar

    def decode(self, latent_vector: torch.Tensor, condition: torch.Tensor | None = None) -> torch.Tensor:
        return self.decoder(latent_vector, condition)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        assert self.unpack_input_condition is not None
        return self._packed_forward(*self.unpack_input_condition(input))
