"""PCA as an ``nn.Module`` (parity: ``fl4health/model_bases/pca.py:12-245``): principal components via (reduced / full)
SVD or randomised low-rank PCA, projection / reconstruction utilities.  All linear algebra stays on the data's device
(cuSOLVER on GPU)."""

from __future__ import annotations

from logging import INFO, WARNING

import torch
from torch import Tensor, nn
from torch.nn.parameter import Parameter

from fl4health_b200.common.logger import log

TWO_D_TENSOR_SHAPE_LENGTH = 2


class PcaModule(nn.Module):
    def __init__(self, low_rank: bool = False, full_svd: bool = False, rank_estimation: int = 6) -> None:
        super().__init__()
        self.low_rank = low_rank
        self.full_svd = full_svd
        self.rank_estimation = rank_estimation
        self.principal_components: Parameter
        self.singular_values: Parameter
        self.data_mean: Tensor

    def forward(self, x: Tensor, center_data: bool) -> tuple[Tensor, Tensor]:
        """(principal components as columns [d, k], singular values [k]) of the (optionally centred) data matrix."""
        x_prime = self.prepare_data_forward(x, center_data=center_data)
        if self.low_rank:
            m, n = x_prime.shape
            if self.rank_estimation > min(m, n):
                log(WARNING, "Estimate of data rank given by user is larger than the actual rank.")
            _, singular_values, components = torch.pca_lowrank(x_prime, q=min(self.rank_estimation, m, n), center=False)
            return components, singular_values
        log(INFO, "Performing full SVD on data matrix." if self.full_svd else "Performing reduced SVD on data matrix.")
        _, singular_values, vh = torch.linalg.svd(x_prime, full_matrices=self.full_svd)
        return vh.T, singular_values

    def maybe_reshape(self, x: Tensor) -> Tensor:
        return torch.squeeze(x.float() if x.dim() == 2 else x.reshape(x.size(0), -1).float())

    def set_data_mean(self, x: Tensor) -> None:
        self.data_mean = torch.mean(x, dim=0)

    def center_data(self, x: Tensor) -> Tensor:
        assert self.data_mean is not None
        return x - self.data_mean

    def prepare_data_forward(self, x: Tensor, center_data: bool) -> Tensor:
        x = self.maybe_reshape(x)
        if center_data:
            self.set_data_mean(x)
            return self.center_data(x)
        mean = torch.mean(x, dim=0)
        assert torch.allclose(torch.zeros_like(mean), mean, atol=1e-6), "data must be centred when center_data=False"
        return x

    def project_lower_dim(self, x: Tensor, k: int | None = None, center_data: bool = False) -> Tensor:
        x_prime = self.maybe_reshape(x)
        if center_data:
            x_prime = self.center_data(x_prime)
        components = self.principal_components[:, :k] if k else self.principal_components
        return x_prime @ components

    def project_back(self, x_lower_dim: Tensor, add_mean: bool = False) -> Tensor:
        low = x_lower_dim.reshape(x_lower_dim.size(0), -1).float()  # (no squeeze: k == 1 must stay a matrix)
        k = low.size(1)
        out = low @ self.principal_components[:, :k].T
        return out + self.data_mean if add_mean else out

    def compute_reconstruction_error(self, x: Tensor, k: int | None, center_data: bool = False) -> float:
        reconstruction = self.project_back(self.project_lower_dim(x, k, center_data=center_data), add_mean=center_data)
        return (torch.linalg.norm(reconstruction - x) ** 2).item() / x.size(0)

    def compute_projection_variance(self, x: Tensor, k: int | None, center_data: bool = False) -> float:
        return (torch.linalg.norm(self.project_lower_dim(x, k, center_data)) ** 2).item()

    def compute_cumulative_explained_variance(self) -> float:
        return torch.sum(self.singular_values**2).item()

    def compute_explained_variance_ratios(self) -> Tensor:
        return (self.singular_values**2) / self.compute_cumulative_explained_variance()

    def set_principal_components(self, principal_components: Tensor, singular_values: Tensor) -> None:
        self.principal_components = Parameter(data=principal_components, requires_grad=False)
        self.singular_values = Parameter(data=singular_values, requires_grad=False)
