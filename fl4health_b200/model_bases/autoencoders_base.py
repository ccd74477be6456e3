"""Auto-encoder wrappers (parity: ``fl4health/model_bases/autoencoders_base.py:8-282``).

The variational models return ONE packed tensor ``[logvar | mu | flattened reconstruction]`` so they fit the
single-tensor prediction contract of the clients; ``preprocessing/autoencoders/loss.py::VaeLoss`` unpacks it.
Both variational flavours share ``_Variational``: encode (with optional conditioning tensors) -> reparameterised sample
-> decode -> pack.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from collections.abc import Callable
from typing import Any

import torch
from torch import nn


def _through(half: str) -> Callable[..., Any]:
    """A method that hands its arguments to ``self.<half>`` (the encoder or the decoder) and returns what it returns:
    ``encode`` / ``decode`` of every auto-encoder flavour, conditioning tensors included."""

    def call(self: nn.Module, *tensors: torch.Tensor | None) -> Any:
        return getattr(self, half)(*tensors)

    call.__name__ = {"encoder": "encode", "decoder": "decode"}[half]
    return call


class AbstractAe(nn.Module, ABC):
    encode = _through("encoder")
    decode = _through("decoder")

    def __init__(self, encoder: nn.Module, decoder: nn.Module) -> None:
        super().__init__()
        self.encoder, self.decoder = encoder, decoder

    @abstractmethod
    def forward(self, input: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError


class BasicAe(AbstractAe):
    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return self.decode(self.encode(input))


class _Variational(AbstractAe):
    """Shared body of the (conditional) variational auto-encoders: ``encode`` returns ``(mu, logvar)``; ``context`` = the
    conditioning tensors (possibly none) handed to both halves of the network."""

    def sampling(self, mu: torch.Tensor, logvar: torch.Tensor) -> torch.Tensor:
        """z = mu + eps * sigma with eps ~ N(0, I) (one fused addcmul)."""
        return torch.addcmul(mu, torch.randn_like(mu), torch.exp(0.5 * logvar))

    def _packed_forward(self, sample: torch.Tensor, *context: torch.Tensor) -> torch.Tensor:
        mu, logvar = self.encode(sample, *context)
        reconstruction = self.decode(self.sampling(mu, logvar), *context)
        return torch.cat((logvar, mu, reconstruction.flatten(start_dim=1)), dim=1)


class VariationalAe(_Variational):
    """``encoder(x) -> (mu, logvar)``; forward returns the packed ``[logvar | mu | recon]`` tensor."""

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return self._packed_forward(input)


class ConditionalVae(_Variational):
    """CVAE: encoder and decoder both receive the condition (``encode(input, condition)``, ``decode(latent, condition)``);
    ``unpack_input_condition`` splits the single input tensor the data loader provides into (input, condition)."""

    def __init__(
        self, encoder: nn.Module, decoder: nn.Module,
        unpack_input_condition: Callable[[torch.Tensor], tuple[torch.Tensor, torch.Tensor]] | None = None,
    ) -> None:
        super().__init__(encoder, decoder)
        self.unpack_input_condition = unpack_input_condition

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        assert self.unpack_input_condition is not None
        return self._packed_forward(*self.unpack_input_condition(input))
