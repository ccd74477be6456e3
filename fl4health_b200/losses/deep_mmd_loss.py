"""Deep-kernel MMD (Liu et al. 2020, "Learning Deep Kernels for Non-Parametric Two-Sample Tests"); parity:
``fl4health/losses/deep_mmd_loss.py:8-329``.

``k_w(a, b) = [(1 - eps) exp(-(|phi(a) - phi(b)|^2 / s_phi)^L) + eps] * exp(-|a - b|^2 / s_q)`` with a small MLP
featurizer ``phi``; in training mode each forward first takes ``optimization_steps`` AdamW steps on the test-power
criterion ``-MMD_u^2 / sigma`` and then returns the MMD estimate under the current kernel.

As in ``mkmmd_loss.py`` every block of pairwise distances comes from ONE Gram GEMM over the concatenated batch.
"""

from __future__ import annotations

import torch


class ModelLatentF(torch.nn.Module):
    """The deep-kernel featurizer: 4 linear layers with Softplus."""

    def __init__(self, x_in_dim: int, hidden_dim: int, x_out_dim: int) -> None:
        super().__init__()
        self.latent = torch.nn.Sequential(
            torch.nn.Linear(x_in_dim, hidden_dim), torch.nn.Softplus(),
            torch.nn.Linear(hidden_dim, hidden_dim), torch.nn.Softplus(),
            torch.nn.Linear(hidden_dim, hidden_dim), torch.nn.Softplus(),
            torch.nn.Linear(hidden_dim, x_out_dim),
        )

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return self.latent(input)


def _pairwise_sq(z: torch.Tensor) -> torch.Tensor:
    sq = (z * z).sum(dim=1)
    return (sq[:, None] + sq[None, :] - 2.0 * (z @ z.t())).clamp_min(0.0)


class DeepMmdLoss(torch.nn.Module):
    def __init__(
        self, device: torch.device, input_size: int, hidden_size: int = 10, output_size: int = 50, lr: float = 0.001,
        is_unbiased: bool = True, gaussian_degree: int = 1, optimization_steps: int = 5,
    ) -> None:
        super().__init__()
        self.device = device
        self.lr = lr
        self.is_unbiased = is_unbiased
        self.gaussian_degree = gaussian_degree
        self.optimization_steps = optimization_steps
        self.featurizer = ModelLatentF(input_size, hidden_size, output_size).to(device)
        self.featurizer.eval()
        # kernel hyper-parameters (plain tensors, optimised alongside the featurizer; same initial values as the reference)
        self.epsilon_opt = torch.log(torch.rand(1, dtype=torch.float64) * 1e-10).to(device)
        self.sigma_q_opt = torch.sqrt(torch.tensor(2.0 * 32 * 32)).to(device)
        self.sigma_phi_opt = torch.sqrt(torch.tensor(0.005)).to(device)
        for t in (self.epsilon_opt, self.sigma_q_opt, self.sigma_phi_opt):
            t.requires_grad = False
        self.optimizer_F = torch.optim.AdamW(
            [*self.featurizer.parameters(), self.epsilon_opt, self.sigma_q_opt, self.sigma_phi_opt], lr=lr
        )
        self.training = False

    def pairwise_distance_squared(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        x_norm, y_norm = (x * x).sum(1).view(-1, 1), (y * y).sum(1).view(1, -1)
        return (x_norm + y_norm - 2.0 * (x @ y.t())).clamp_min(0.0)

    def h1_mean_var_gram(
        self, k_x: torch.Tensor, k_y: torch.Tensor, k_xy: torch.Tensor, is_var_computed: bool
    ) -> tuple[torch.Tensor, torch.Tensor | None]:
        nx, ny = k_x.shape[0], k_y.shape[0]
        if self.is_unbiased:
            xx = (k_x.sum() - torch.diagonal(k_x).sum()) / (nx * (nx - 1))
            yy = (k_y.sum() - torch.diagonal(k_y).sum()) / (ny * (ny - 1))
            xy = (k_xy.sum() - torch.diagonal(k_xy).sum()) / (nx * (ny - 1))
        else:
            xx, yy, xy = k_x.sum() / (nx * nx), k_y.sum() / (ny * ny), k_xy.sum() / (nx * ny)
        mmd2 = xx - 2 * xy + yy
        if not is_var_computed:
            return mmd2, None
        h_ij = k_x + k_y - k_xy - k_xy.t()
        row = h_ij.sum(1)
        variance = (4.0 / ny**3) * torch.dot(row, row) - (4.0 / nx**4) * (h_ij.sum() ** 2) + 1e-8
        return mmd2, variance

    def mmdu(
        self, features: torch.Tensor, len_s: int, features_org: torch.Tensor, sigma_q: torch.Tensor,
        sigma_phi: torch.Tensor, epsilon: torch.Tensor, is_smooth: bool = True, is_var_computed: bool = True,
    ) -> tuple[torch.Tensor, torch.Tensor | None]:
        d_feat = _pairwise_sq(features)  # one GEMM: xx, yy and xy blocks
        if is_smooth:
            d_org = _pairwise_sq(features_org)
            base = torch.exp(-d_org / sigma_q)
            kernel = ((1 - epsilon) * torch.exp(-((d_feat / sigma_phi) ** self.gaussian_degree)) + epsilon) * base
        else:
            kernel = torch.exp(-d_feat / sigma_phi)
        kernel = kernel.to(features.dtype)
        return self.h1_mean_var_gram(kernel[:len_s, :len_s], kernel[len_s:, len_s:], kernel[:len_s, len_s:], is_var_computed)

    def _kernel_params(self) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        epsilon = torch.sigmoid(self.epsilon_opt).to(torch.float32)
        return epsilon, self.sigma_q_opt**2, self.sigma_phi_opt**2

    def _set_kernel_grad(self, flag: bool) -> None:
        for t in (self.epsilon_opt, self.sigma_q_opt, self.sigma_phi_opt):
            t.requires_grad = flag

    def train_kernel(self, x: torch.Tensor, y: torch.Tensor) -> None:
        """One AdamW step maximising the test-power proxy ``MMD_u^2 / sqrt(var)``."""
        self.featurizer.train()
        self._set_kernel_grad(True)
        features = torch.cat([x, y[torch.randperm(y.size(0), device=y.device)]], 0)
        self.optimizer_F.zero_grad()
        epsilon, sigma_q, sigma_phi = self._kernel_params()
        mmd, var = self.mmdu(self.featurizer(features), x.shape[0], features.view(features.shape[0], -1), sigma_q, sigma_phi,
                             epsilon, is_var_computed=True)
        assert var is not None
        (-mmd / torch.sqrt(var)).backward()
        self.optimizer_F.step()

    def compute_kernel(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        self.featurizer.eval()
        self._set_kernel_grad(False)
        features = torch.cat([x, y], 0)
        epsilon, sigma_q, sigma_phi = self._kernel_params()
        mmd, _ = self.mmdu(self.featurizer(features), x.shape[0], features.view(features.shape[0], -1), sigma_q, sigma_phi,
                           epsilon, is_var_computed=False)
        return mmd

    def forward(self, x_s: torch.Tensor, x_t: torch.Tensor) -> torch.Tensor:
        if self.training:
            with torch.enable_grad():
                for _ in range(self.optimization_steps):
                    self.train_kernel(x_s.clone().detach(), x_t.clone().detach())
        return self.compute_kernel(x_s, x_t)
