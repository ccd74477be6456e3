"""Synthetic(alpha, beta) data of the FedProx paper (Li et al. 2020, Sec. 5.1 / App. C.1); parity:
``fl4health/utils/data_generation.py:12-353``.

Client k draws inputs ``x ~ N(v_k, Sigma)`` with ``Sigma_jj = j^-1.2`` and ``v_k ~ N(B_k, 1)``, ``B_k ~ N(0, beta)``; labels
are ``argmax softmax((W_k x + b_k) / T)`` with ``W_k, b_k ~ N(u_k, 1)``, ``u_k ~ N(0, alpha)`` (optionally through a hidden
layer).  The IID variant shares one ``(W, b)`` and a centred input distribution.  Inputs are sampled with one
``randn * sqrt(diag)`` instead of a full ``MultivariateNormal`` (the covariance is diagonal).
"""

from __future__ import annotations

from abc import ABC, abstractmethod

import torch
import torch.nn.functional as F

from fl4health_b200.utils.dataset import TensorDataset


class SyntheticFedProxDataset(ABC):
    def __init__(
        self, num_clients: int, temperature: float = 1.0, input_dim: int = 60, output_dim: int = 10,
        hidden_dim: int | None = None, samples_per_client: int = 1000,
    ) -> None:
        self.num_clients = num_clients
        self.temperature = temperature
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.hidden_dim = hidden_dim
        self.samples_per_client = samples_per_client
        self.input_covariance = self.construct_covariance_matrix()

    def construct_covariance_matrix(self) -> torch.Tensor:
        return torch.diag(torch.arange(1, self.input_dim + 1, dtype=torch.float32).pow(-1.2))

    def _sample_inputs(self, mean: torch.Tensor) -> torch.Tensor:
        std = torch.sqrt(torch.diagonal(self.input_covariance))
        return mean + torch.randn(self.samples_per_client, self.input_dim) * std

    def _hard_labels(self, logits: torch.Tensor) -> torch.Tensor:
        return F.one_hot(torch.argmax(F.softmax(logits, dim=1), dim=1), num_classes=self.output_dim)

    def one_layer_map_inputs_to_outputs(self, x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        """x: [n, in], w: [out, in], b: [out, 1] -> one-hot labels [n, out]."""
        return self._hard_labels((x @ w.T + b.T) / self.temperature)

    def two_layer_map_inputs_to_outputs(
        self, x: torch.Tensor, w_1: torch.Tensor, b_1: torch.Tensor, w_2: torch.Tensor, b_2: torch.Tensor
    ) -> torch.Tensor:
        latent = (x @ w_1.T + b_1.T) / self.temperature
        return self._hard_labels(latent @ w_2.T + b_2.T)

    def generate(self) -> list[TensorDataset]:
        client_tensors = self.generate_client_tensors()
        assert len(client_tensors) == self.num_clients, (
            "The tensors returned by generate_client_tensors should have the same length as self.num_clients")
        return [TensorDataset(x, y) for x, y in client_tensors]

    @abstractmethod
    def generate_client_tensors(self) -> list[tuple[torch.Tensor, torch.Tensor]]:
        raise NotImplementedError


class SyntheticNonIidFedProxDataset(SyntheticFedProxDataset):
    def __init__(
        self, num_clients: int, alpha: float, beta: float, temperature: float = 1.0, input_dim: int = 60,
        output_dim: int = 10, hidden_dim: int | None = None, samples_per_client: int = 1000,
    ) -> None:
        super().__init__(num_clients, temperature, input_dim, output_dim, hidden_dim, samples_per_client)
        self.two_layer_generation = hidden_dim is not None
        self.alpha = alpha
        self.beta = beta

    def get_input_output_tensors(self, mu: list[float], v: torch.Tensor, sigma: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:  # noqa: ARG002
        x = self._sample_inputs(v)
        if not self.two_layer_generation:
            w = mu[0] + torch.randn(self.output_dim, self.input_dim)
            b = mu[0] + torch.randn(self.output_dim, 1)
            return x, self.one_layer_map_inputs_to_outputs(x, w, b)
        assert self.hidden_dim is not None
        w_1, b_1 = mu[0] + torch.randn(self.hidden_dim, self.input_dim), mu[0] + torch.randn(self.hidden_dim, 1)
        w_2, b_2 = mu[1] + torch.randn(self.output_dim, self.hidden_dim), mu[1] + torch.randn(self.output_dim, 1)
        return x, self.two_layer_map_inputs_to_outputs(x, w_1, b_1, w_2, b_2)

    def generate_client_tensors(self) -> list[tuple[torch.Tensor, torch.Tensor]]:
        out = []
        for _ in range(self.num_clients):
            b_k = torch.randn(1) * self.beta
            input_means = b_k + torch.randn(self.input_dim)  # v_k
            means = [float(torch.randn(1) * self.alpha)]  # u_k
            if self.two_layer_generation:
                means.append(float(torch.randn(1) * self.alpha))
            out.append(self.get_input_output_tensors(means, input_means, self.input_covariance))
        return out


class SyntheticIidFedProxDataset(SyntheticFedProxDataset):
    def __init__(self, num_clients: int, temperature: float = 1.0, input_dim: int = 60, output_dim: int = 10,
                 samples_per_client: int = 1000) -> None:
        super().__init__(num_clients, temperature, input_dim, output_dim, None, samples_per_client)
        self.w = torch.randn(self.output_dim, self.input_dim)
        self.b = torch.randn(self.output_dim, 1)

    def get_input_output_tensors(self) -> tuple[torch.Tensor, torch.Tensor]:
        x = self._sample_inputs(torch.zeros(self.input_dim))
        return x, self.one_layer_map_inputs_to_outputs(x, self.w, self.b)

    def generate_client_tensors(self) -> list[tuple[torch.Tensor, torch.Tensor]]:
        return [self.get_input_output_tensors() for _ in range(self.num_clients)]
