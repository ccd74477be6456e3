"""bench.py --config variants (benchmarks/fl_variants.py): every algorithm of the BASELINE configurations builds through
the public client / strategy / server API and runs federated rounds on a single CPU rank."""

import sys
from pathlib import Path

import pytest
import torch
from torch import nn

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "benchmarks"))

import fl_variants  # noqa: E402

from fl4health_b200.engine.data import BatchedTensorLoader  # noqa: E402
from fl4health_b200.engine.options import EngineOptions  # noqa: E402
from fl4health_b200.parallel.spmd import SpmdContext, build_spmd_federation  # noqa: E402
from fl4health_b200.utils.dataset import TensorDataset  # noqa: E402


def _data(n: int, seed: int) -> TensorDataset:
    gen = torch.Generator().manual_seed(seed)
    return TensorDataset(torch.randn(n, 3, 32, 32, generator=gen), torch.randint(0, 10, (n,), generator=gen))


class _Hooks:
    def get_data_loaders(self, config):  # noqa: ANN001, ANN202
        bs = int(config["batch_size"])
        return (BatchedTensorLoader(_data(4 * bs, 1), bs, shuffle=True, drop_last=True, device=self.device),
                BatchedTensorLoader(_data(bs, 2), bs, device=self.device))

    def get_criterion(self, config):  # noqa: ANN001, ANN202
        return nn.CrossEntropyLoss()

    def get_optimizer(self, config):  # noqa: ANN001, ANN202
        return torch.optim.SGD(self.model.parameters(), lr=0.01, momentum=0.9)


@pytest.mark.parametrize("variant", [v for group in fl_variants.GROUPS.values() for v in group])
def test_variant_runs_two_rounds(variant: str, monkeypatch) -> None:
    monkeypatch.setenv("WORLD_SIZE", "1")
    monkeypatch.setenv("RANK", "0")
    ctx = SpmdContext()
    engine = EngineOptions(arena=True, fused_optimizer=False, cuda_graphs=False)
    client, server = fl_variants.build(variant, _Hooks, ctx, engine, rounds=2, local_steps=2, batch_size=8)
    build_spmd_federation(ctx, server, client, fused=False)
    history, _ = server.fit(num_rounds=2)
    losses = [loss for _, loss in history.losses_distributed]
    assert len(losses) == 2 and all(loss == loss for loss in losses)
