"""Warm-start surgery, PCA / auto-encoder preprocessing transforms, VAE loss."""

import json
from pathlib import Path

import pytest
import torch
from torch import nn

from fl4health_b200.model_bases.autoencoders_base import BasicAe, ConditionalVae, VariationalAe
from fl4health_b200.model_bases.pca import PcaModule
from fl4health_b200.parallel.arena import attach_arena
from fl4health_b200.preprocessing.autoencoders.dim_reduction import (
    AeProcessor,
    CvaeFixedConditionProcessor,
    CvaeVariableConditionProcessor,
    VaeProcessor,
)
from fl4health_b200.preprocessing.autoencoders.loss import VaeLoss
from fl4health_b200.preprocessing.pca_preprocessor import PcaPreprocessor
from fl4health_b200.preprocessing.warmed_up_module import WarmedUpModule
from fl4health_b200.utils.dataset import TensorDataset


class Small(nn.Module):
    def __init__(self, out: int = 4) -> None:
        super().__init__()
        self.features = nn.Sequential(nn.Linear(6, 8), nn.ReLU(), nn.Linear(8, 8))
        self.head = nn.Linear(8, out)


class Renamed(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.encoder = nn.Sequential(nn.Linear(6, 8), nn.ReLU(), nn.Linear(8, 8))
        self.classifier = nn.Linear(8, 3)  # different size: must NOT be loaded


@pytest.mark.parametrize("arena", [False, True])
def test_warm_up_direct_and_mapped(tmp_path: Path, arena: bool) -> None:
    torch.manual_seed(0)
    pretrained = Small()
    target = Small()
    if arena:
        attach_arena(target)
    WarmedUpModule(pretrained_model=pretrained).load_from_pretrained(target)
    assert all(torch.equal(a, b) for a, b in zip(pretrained.state_dict().values(), target.state_dict().values()))
    mapping = tmp_path / "map.json"
    mapping.write_text(json.dumps({"encoder": "features", "classifier": "head"}))
    renamed = Renamed()
    before = renamed.classifier.weight.detach().clone()
    torch.save(pretrained, tmp_path / "pre.pt")
    warm = WarmedUpModule(pretrained_model_path=tmp_path / "pre.pt", weights_mapping_path=mapping)
    assert warm.get_matching_component("encoder.2.bias") == "features.2.bias" and warm.get_matching_component("other.w") is None
    warm.load_from_pretrained(renamed)
    assert torch.equal(renamed.encoder[0].weight, pretrained.features[0].weight)
    assert torch.equal(renamed.classifier.weight, before)  # shape mismatch -> untouched
    with pytest.raises(AssertionError):
        WarmedUpModule()


def test_pca_preprocessor(tmp_path: Path) -> None:
    torch.manual_seed(1)
    x = torch.randn(50, 6) * torch.tensor([5.0, 3.0, 1.0, 0.1, 0.1, 0.1])
    pca = PcaModule()
    comps, vals = pca(x, center_data=True)
    pca.set_principal_components(comps, vals)
    torch.save(pca, tmp_path / "pca.pt")
    ds = PcaPreprocessor(tmp_path / "pca.pt").reduce_dimension(2, TensorDataset(x, torch.zeros(50)))
    assert ds[0][0].shape == (2,)
    batch, _ = ds.get_batch(torch.arange(5))
    assert batch.shape == (5, 2) and torch.allclose(batch[0], ds[0][0], atol=1e-5)


class _Enc(nn.Module):
    def __init__(self, cond: int = 0) -> None:
        super().__init__()
        self.mu, self.logvar = nn.Linear(6 + cond, 3), nn.Linear(6 + cond, 3)

    def forward(self, x, condition=None):
        if condition is not None:
            x = torch.cat((x, condition.expand(x.shape[0], -1) if condition.dim() == 1 else condition), dim=-1)
        return self.mu(x), self.logvar(x)


class _Dec(nn.Module):
    def __init__(self, cond: int = 0) -> None:
        super().__init__()
        self.fc = nn.Linear(3 + cond, 6)

    def forward(self, z, condition=None):
        if condition is not None:
            z = torch.cat((z, condition), dim=-1)
        return self.fc(z)


def test_autoencoder_processors_and_vae_loss(tmp_path: Path) -> None:
    torch.manual_seed(2)
    cpu = torch.device("cpu")
    x = torch.randn(5, 6)
    torch.save(BasicAe(nn.Linear(6, 3), nn.Linear(3, 6)), tmp_path / "ae.pt")
    assert AeProcessor(tmp_path / "ae.pt", cpu)(x).shape == (5, 3)
    vae = VariationalAe(_Enc(), _Dec())
    torch.save(vae, tmp_path / "vae.pt")
    assert VaeProcessor(tmp_path / "vae.pt", cpu)(x).shape == (5, 6)
    assert VaeProcessor(tmp_path / "vae.pt", cpu, return_mu_only=True)(x[0]).shape == (3,)
    cvae = ConditionalVae(_Enc(2), _Dec(2), unpack_input_condition=lambda t: (t[:, :6], t[:, 6:]))
    torch.save(ConditionalVae(_Enc(2), _Dec(2)), tmp_path / "cvae.pt")
    fixed = CvaeFixedConditionProcessor(tmp_path / "cvae.pt", torch.tensor([1.0, 0.0]), cpu)
    assert fixed(x).shape == (5, 6)
    variable = CvaeVariableConditionProcessor(tmp_path / "cvae.pt", cpu, return_mu_only=True)
    assert variable(x, torch.rand(5, 2)).shape == (5, 3)
    packed = vae(x)
    assert packed.shape == (5, 3 + 3 + 6)
    loss = VaeLoss(3, nn.MSELoss(reduction="sum"))
    recon, mu, logvar = loss.unpack_model_output(packed)
    expected = ((recon - x) ** 2).sum() - 0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp())
    assert loss(packed, x).item() == pytest.approx(expected.item(), rel=1e-6)
    assert loss.standard_normal_kl_divergence_loss(torch.zeros(2, 3), torch.zeros(2, 3)).item() == 0.0
    assert cvae(torch.cat((x, torch.rand(5, 2)), dim=1)).shape == (5, 12)
