"""Build real clients around tiny models without a server (the reference's "client integration" tier,
tests/clients/fixtures.py:49-216): inject model/optimizer/loaders and mark the client initialised."""

from __future__ import annotations

from pathlib import Path

import torch
from torch import nn

from fl4health_b200.engine.data import BatchedTensorLoader
from fl4health_b200.engine.options import EngineOptions
from fl4health_b200.metrics import Accuracy
from fl4health_b200.utils.dataset import TensorDataset


class SmallMlp(nn.Module):
    def __init__(self, d_in: int = 8, hidden: int = 6, classes: int = 3) -> None:
        super().__init__()
        self.fc1 = nn.Linear(d_in, hidden)
        self.fc2 = nn.Linear(hidden, classes)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.fc2(torch.relu(self.fc1(x)))


class LinearTransform(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.linear = nn.Linear(2, 3, bias=False)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.linear(x)


def toy_loaders(d_in: int = 8, classes: int = 3, n: int = 64, batch: int = 16, seed: int = 0):  # noqa: ANN201
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d_in, generator=gen)
    y = torch.randint(0, classes, (n,), generator=gen)
    return BatchedTensorLoader(TensorDataset(x, y), batch), BatchedTensorLoader(TensorDataset(x.clone(), y.clone()), batch)


def build_client(client_cls, model: nn.Module, *, arena: bool = True, lr: float = 0.1, optimizer: str = "sgd",  # noqa: ANN001, ANN201
                 d_in: int = 8, classes: int = 3, **client_kwargs):
    """Instantiate ``client_cls`` and wire the pieces ``setup_client`` would, using the engine's own placement."""
    engine = EngineOptions(arena=arena, fused_optimizer=arena)
    client = client_cls(Path("."), [Accuracy()], torch.device("cpu"), client_name="fixture", engine_options=engine,
                        **client_kwargs)

    def get_model(config):  # noqa: ANN001, ANN202
        return model

    def get_optimizer(config):  # noqa: ANN001, ANN202
        params = client.model.parameters()
        return torch.optim.SGD(params, lr=lr) if optimizer == "sgd" else torch.optim.AdamW(params, lr=lr)

    client.get_model = get_model
    client.get_optimizer = get_optimizer
    client.get_data_loaders = lambda config: toy_loaders(d_in, classes)
    client.get_criterion = lambda config: nn.CrossEntropyLoss()
    client.setup_client({"current_server_round": 1, "batch_size": 16})
    return client
