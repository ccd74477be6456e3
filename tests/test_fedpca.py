"""Federated PCA: SVD- and QR-merging recover the principal subspace of the pooled data; PcaModule utilities."""

from pathlib import Path

import pytest
import torch

from fl4health_b200.clients.fed_pca_client import FedPCAClient
from fl4health_b200.engine.data import BatchedTensorLoader
from fl4health_b200.model_bases.pca import PcaModule
from fl4health_b200.servers.base_server import FlServer
from fl4health_b200.servers.client_manager import SimpleClientManager
from fl4health_b200.simulation import run_simulation
from fl4health_b200.strategies.fedpca import FedPCA
from fl4health_b200.utils.dataset import TensorDataset


def _data(seed: int, n: int = 200, d: int = 10) -> torch.Tensor:
    gen = torch.Generator().manual_seed(seed)
    basis = torch.linalg.qr(torch.randn(d, d, generator=torch.Generator().manual_seed(0)))[0]
    scales = torch.tensor([10.0, 6.0, 3.0] + [0.1] * (d - 3))
    x = (torch.randn(n, d, generator=gen) * scales) @ basis.T
    return x - x.mean(0)


def test_pca_module_reconstruction() -> None:
    x = _data(1)
    pca = PcaModule()
    comps, vals = pca(x, center_data=True)
    pca.set_principal_components(comps, vals)
    assert pca.compute_reconstruction_error(x, None) < 1e-6
    assert pca.compute_reconstruction_error(x, 3) < 0.2 and pca.compute_reconstruction_error(x, 1) > 1.0
    assert pca.compute_explained_variance_ratios().sum().item() == pytest.approx(1.0, abs=1e-5)
    low = PcaModule(low_rank=True, rank_estimation=3)
    comps_l, _ = low(x, center_data=True)
    overlap = torch.linalg.svdvals(comps[:, :3].T @ comps_l[:, :3])
    assert overlap.min() > 0.99


@pytest.mark.parametrize("svd_merging", [True, False])
def test_subspace_merging_matches_pooled_svd(svd_merging: bool) -> None:
    parts = [_data(s) for s in (1, 2, 3)]
    pooled = torch.cat(parts)
    _, s_pool, vh = torch.linalg.svd(pooled.double(), full_matrices=False)
    client_u, client_s = [], []
    for p in parts:
        _, s, v = torch.linalg.svd(p.double(), full_matrices=False)
        client_u.append(v.T)
        client_s.append(s)
    strategy = FedPCA(svd_merging=svd_merging)
    merge = strategy.merge_subspaces_svd if svd_merging else strategy.merge_subspaces_qr
    u, s = merge(client_u, client_s)
    assert torch.allclose(s[:3], s_pool[:3], rtol=1e-6)
    overlap = torch.linalg.svdvals(u[:, :3].T @ vh[:3].T)
    assert overlap.min() > 1 - 1e-8


def test_fedpca_end_to_end(tmp_path: Path) -> None:
    class Client(FedPCAClient):
        def __init__(self, idx: int) -> None:
            super().__init__(Path("."), torch.device("cpu"), tmp_path, client_name=f"p{idx}")
            self.idx = idx

        def get_data_loaders(self, config):
            x = _data(10 + self.idx)
            ds = TensorDataset(x, torch.zeros(len(x)))
            return BatchedTensorLoader(ds, 64), BatchedTensorLoader(TensorDataset(_data(20 + self.idx, n=64), torch.zeros(64)), 64)

    def cfg(r):
        return {"current_server_round": r, "low_rank": False, "full_svd": False, "rank_estimation": 3, "center_data": True,
                "num_components_eval": 3}

    strategy = FedPCA(min_fit_clients=2, min_evaluate_clients=2, min_available_clients=2, on_fit_config_fn=cfg,
                      on_evaluate_config_fn=cfg)
    server = FlServer(SimpleClientManager(), {"n_server_rounds": 1}, strategy, on_init_parameters_config_fn=cfg)
    history = run_simulation(server, [Client(0), Client(1)], 1)
    assert history.losses_distributed[0][1] < 0.5  # reconstruction error with the top-3 merged components
    assert (tmp_path / "client_p0_pca.pt").exists()
