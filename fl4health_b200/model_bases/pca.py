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
    """``forward`` decomposes a data matrix; the ``project_*`` / ``compute_*`` utilities work with whatever components
    were installed by ``set_principal_components`` (locally computed or merged by the FedPCA strategy)."""

    def __init__(self, low_rank: bool = False, full_svd: bool = False, rank_estimation: int = 6) -> None:
        super().__init__()
        self.low_rank, self.full_svd, self.rank_estimation = low_rank, full_svd, rank_estimation
        self.principal_components: Parameter
        self.singular_values: Parameter
        self.data_mean: Tensor

    # ---- data preparation -------------------------------------------------------------------------------------
    def maybe_reshape(self, x: Tensor) -> Tensor:
        """Samples as rows of a float matrix (anything beyond the first axis is flattened)."""
        matrix = x if x.dim() == TWO_D_TENSOR_SHAPE_LENGTH else x.reshape(x.size(0), -1)
        return torch.squeeze(matrix.float())

    def set_data_mean(self, x: Tensor) -> None:
        self.data_mean = x.mean(dim=0)

    def center_data(self, x: Tensor) -> Tensor:
        assert self.data_mean is not None
        return x - self.data_mean

    def prepare_data_forward(self, x: Tensor, center_data: bool) -> Tensor:
        matrix = self.maybe_reshape(x)
        if not center_data:
            column_means = matrix.mean(dim=0)
            assert torch.allclose(column_means, torch.zeros_like(column_means), atol=1e-6), \
                "data must be centred when center_data=False"
            return matrix
        self.set_data_mean(matrix)
        return self.center_data(matrix)

    # ---- decomposition ----------------------------------------------------------------------------------------
    def _randomized(self, matrix: Tensor) -> tuple[Tensor, Tensor]:
        smallest_side = min(matrix.shape)
        if self.rank_estimation > smallest_side:
            log(WARNING, "Estimate of data rank given by user is larger than the actual rank.")
        _, singular_values, components = torch.pca_lowrank(matrix, q=min(self.rank_estimation, smallest_side), center=False)
        return components, singular_values

    def _exact(self, matrix: Tensor) -> tuple[Tensor, Tensor]:
        log(INFO, "Performing full SVD on data matrix." if self.full_svd else "Performing reduced SVD on data matrix.")
        _, singular_values, right_vectors = torch.linalg.svd(matrix, full_matrices=self.full_svd)
        return right_vectors.T, singular_values

    def forward(self, x: Tensor, center_data: bool) -> tuple[Tensor, Tensor]:
        """(principal components as columns [d, k], singular values [k]) of the (optionally centred) data matrix."""
        matrix = self.prepare_data_forward(x, center_data=center_data)
        return self._randomized(matrix) if self.low_rank else self._exact(matrix)

    def set_principal_components(self, principal_components: Tensor, singular_values: Tensor) -> None:
        self.principal_components = Parameter(data=principal_components, requires_grad=False)
        self.singular_values = Parameter(data=singular_values, requires_grad=False)

    # ---- projections ------------------------------------------------------------------------------------------
    def _basis(self, k: int | None) -> Tensor:
        return self.principal_components[:, :k] if k else self.principal_components

    def project_lower_dim(self, x: Tensor, k: int | None = None, center_data: bool = False) -> Tensor:
        matrix = self.maybe_reshape(x)
        return (self.center_data(matrix) if center_data else matrix) @ self._basis(k)

    def project_back(self, x_lower_dim: Tensor, add_mean: bool = False) -> Tensor:
        codes = x_lower_dim.reshape(x_lower_dim.size(0), -1).float()  # (no squeeze: k == 1 must stay a matrix)
        restored = codes @ self._basis(codes.size(1)).T
        return restored + self.data_mean if add_mean else restored

    # ---- diagnostics ------------------------------------------------------------------------------------------
    def compute_reconstruction_error(self, x: Tensor, k: int | None, center_data: bool = False) -> float:
        round_trip = self.project_back(self.project_lower_dim(x, k, center_data=center_data), add_mean=center_data)
        return torch.linalg.norm(round_trip - x).square().item() / x.size(0)

    def compute_projection_variance(self, x: Tensor, k: int | None, center_data: bool = False) -> float:
        return torch.linalg.norm(self.project_lower_dim(x, k, center_data)).square().item()

    def compute_cumulative_explained_variance(self) -> float:
        return self.singular_values.square().sum().item()

    def compute_explained_variance_ratios(self) -> Tensor:
        return self.singular_values.square() / self.compute_cumulative_explained_variance()
