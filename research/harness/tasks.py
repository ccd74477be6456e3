"""Tasks = data + model zoo.  Every task exposes the same small surface so that any method in
``research.harness.methods`` can be run on it:

* ``client_data(client_index, spec) -> (train, val, test)`` ``TensorDataset`` triples (real partitions when they were
  pre-processed to ``spec.data_dir``; otherwise seeded synthetic data of the same shape, so the harness runs anywhere);
* ``features()`` / ``head()`` / ``parallel_head()`` — building blocks for plain, sequentially split and parallel models.
"""

from __future__ import annotations

from collections.abc import Callable
from dataclasses import dataclass, field
from pathlib import Path
from typing import TYPE_CHECKING, Any

import torch
from torch import nn

from fl4health_b200.metrics.metrics import Accuracy
from fl4health_b200.utils.dataset import TensorDataset

if TYPE_CHECKING:
    from research.harness.experiment import ExperimentSpec

Triple = tuple[TensorDataset, TensorDataset, TensorDataset]


@dataclass
class Task:
    name: str
    n_clients: int
    class_num: int
    features: Callable[[], nn.Module]
    head: Callable[[], nn.Module]
    parallel_head: Callable[[], nn.Module]
    client_data: Callable[[int, ExperimentSpec], Triple]
    criterion: Callable[[], nn.Module] = nn.CrossEntropyLoss
    metrics: Callable[[], list[Any]] = field(default=lambda: [Accuracy()])
    feature_layers: tuple[str, ...] = ("features",)

    def plain(self) -> nn.Module:
        from research.cifar10.model import SplitClassifier

        return SplitClassifier(self.features(), self.head())


TASKS: dict[str, Callable[..., Task]] = {}


def task(name: str) -> Callable[[Callable[..., Task]], Callable[..., Task]]:
    def register(builder: Callable[..., Task]) -> Callable[..., Task]:
        TASKS[name] = builder
        return builder

    return register


def split_three_ways(dataset: TensorDataset, seed: int, val_fraction: float = 0.15, test_fraction: float = 0.15) -> Triple:
    assert dataset.targets is not None
    order = torch.randperm(len(dataset.data), generator=torch.Generator().manual_seed(seed))
    n_val, n_test = int(len(order) * val_fraction), int(len(order) * test_fraction)
    cut = lambda idx: TensorDataset(dataset.data[idx], dataset.targets[idx])  # noqa: E731
    return cut(order[n_val + n_test:]), cut(order[:n_val]), cut(order[n_val:n_val + n_test])


def _label_skewed_synthetic(shape: tuple[int, ...], class_num: int, n: int, beta: float, seed: int) -> TensorDataset:
    """Class-separable Gaussian blobs whose label marginal is drawn from Dirichlet(beta) — a stand-in for one client's
    shard of a Dirichlet partition."""
    gen = torch.Generator().manual_seed(seed)
    prior = torch.distributions.Dirichlet(torch.full((class_num,), float(beta))).sample() if beta > 0 else torch.full((class_num,), 1.0 / class_num)
    targets = torch.multinomial(prior, n, replacement=True, generator=gen)
    centres = torch.randn(class_num, *shape, generator=torch.Generator().manual_seed(1234)) * 0.35  # shared by all clients
    data = centres[targets] + torch.randn(n, *shape, generator=gen) * 0.6
    return TensorDataset(data, targets)


def _load_partition(directory: Path, client_index: int) -> Triple | None:
    files = [directory / f"client_{client_index}_{part}.pt" for part in ("train", "val", "test")]
    if not all(f.exists() for f in files):
        return None
    out = []
    for f in files:
        blob = torch.load(f, weights_only=True)
        out.append(TensorDataset(blob["data"], blob["targets"]))
    return out[0], out[1], out[2]


# ------------------------------------------------------------------------------------------------------------ CIFAR-10
@task("cifar10")
def cifar10(n_clients: int = 5, hidden: int = 2048, use_bn: bool = True) -> Task:
    """pFL benchmark of ``research/cifar10``: Dirichlet(β)-partitioned CIFAR-10 (``research.cifar10.preprocess``) and
    the 8.47 M-parameter ConvNet."""
    from research.cifar10.model import ConcatClassifier, ConvFeatures, MlpClassifier, feature_dim

    def client_data(client_index: int, spec: ExperimentSpec) -> Triple:
        real = _load_partition(Path(spec.data_dir) / f"beta_{spec.heterogeneity}" / f"seed_{spec.data_seed}", client_index)
        if real is not None:
            return real
        torch.manual_seed(spec.data_seed * 1000 + client_index)
        full = _label_skewed_synthetic((3, 32, 32), 10, spec.samples_per_client, spec.heterogeneity, spec.data_seed * 1000 + client_index)
        return split_three_ways(full, spec.data_seed + client_index)

    return Task("cifar10", n_clients, 10, lambda: ConvFeatures(3, use_bn), lambda: MlpClassifier(feature_dim(), hidden, 10),
                lambda: ConcatClassifier(feature_dim(), hidden, 10), client_data)


# ----------------------------------------------------------------------------------------------- synthetic (FedProx paper)
@task("synthetic")
def synthetic(n_clients: int = 8, alpha: float = 0.5, beta: float = 0.5, input_dim: int = 60, hidden: int = 64) -> Task:
    """``research/synthetic_data``: the FedProx-paper generator (``SyntheticNonIidFedProxDataset``; α controls model
    heterogeneity across clients, β the input-distribution heterogeneity) and a two-layer MLP."""
    from fl4health_b200.utils.data_generation import SyntheticNonIidFedProxDataset
    from research.cifar10.model import ConcatClassifier

    cache: dict[int, list[TensorDataset]] = {}

    def client_data(client_index: int, spec: ExperimentSpec) -> Triple:
        if spec.data_seed not in cache:
            torch.manual_seed(spec.data_seed)
            generator = SyntheticNonIidFedProxDataset(n_clients, alpha, beta, input_dim=input_dim, output_dim=10,
                                                      samples_per_client=spec.samples_per_client)
            shards = generator.generate()
            cache[spec.data_seed] = [TensorDataset(s.data, s.targets.argmax(dim=1) if s.targets.dim() > 1 else s.targets) for s in shards]
        return split_three_ways(cache[spec.data_seed][client_index], spec.data_seed + client_index)

    return Task("synthetic", n_clients, 10, lambda: nn.Sequential(nn.Linear(input_dim, hidden), nn.ReLU()),
                lambda: nn.Linear(hidden, 10), lambda: ConcatClassifier(hidden, hidden, 10), client_data)


# ---------------------------------------------------------------------------------------------------------------- RxRx1
class _ResNetFeatures(nn.Module):
    def __init__(self, in_channels: int) -> None:
        super().__init__()
        from fl4health_b200.models.resnet import ResNet18

        self.net = ResNet18(num_classes=1, in_channels=in_channels, imagenet_stem=True)
        self.net.fc = nn.Identity()  # type: ignore[assignment]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.net.forward_features(x)


@task("rxrx1")
def rxrx1(n_clients: int = 4, class_num: int = 50, image_size: int = 64) -> Task:
    """``research/rxrx1``: one client per microscopy site, ResNet-18 backbone (the reference fine-tunes torchvision's
    pretrained one; there is no weight download here, so the backbone starts from He initialisation).  Real data is
    read from ``data_dir/rxrx1/client_{i}_{train,val,test}.pt`` tensors produced by ``research.rxrx1.preprocess``."""
    from research.cifar10.model import ConcatClassifier

    def client_data(client_index: int, spec: ExperimentSpec) -> Triple:
        real = _load_partition(Path(spec.data_dir) / "rxrx1", client_index)
        if real is not None:
            return real
        full = _label_skewed_synthetic((3, image_size, image_size), class_num, spec.samples_per_client, 0.0, spec.data_seed * 77 + client_index)
        full.data += 0.3 * client_index  # a site-specific intensity shift: the covariate shift the benchmark studies
        return split_three_ways(full, spec.data_seed + client_index)

    return Task("rxrx1", n_clients, class_num, lambda: _ResNetFeatures(3), lambda: nn.Linear(512, class_num),
                lambda: ConcatClassifier(512, 512, class_num), client_data, feature_layers=("features",))


# -------------------------------------------------------------------------------------------------------------- AG News
class _BertFeatures(nn.Module):
    def __init__(self, vocab_size: int, tiny: bool) -> None:
        super().__init__()
        from fl4health_b200.models.bert import BertConfig, BertEncoder

        cfg = BertConfig.tiny(vocab_size) if tiny else BertConfig(vocab_size=vocab_size)
        self.hidden_size = cfg.hidden_size
        self.bert = BertEncoder(cfg)
        self.pooler = nn.Linear(cfg.hidden_size, cfg.hidden_size)

    def forward(self, input_ids: torch.Tensor) -> torch.Tensor:
        return torch.tanh(self.pooler(self.bert(input_ids, (input_ids != 0).long(), None)[:, 0]))


@task("ag_news")
def ag_news(n_clients: int = 4, vocab_size: int = 1000, seq_len: int = 32, tiny: bool = True) -> Task:
    """``research/ag_news``: 4-way news-topic classification with a BERT encoder, the test bed of the partial-exchange
    studies (dynamic layer / sparse COO exchange).  Synthetic token sequences (topic-specific vocabulary bands) stand in
    for the tokenised corpus unless ``data_dir/ag_news/client_*`` tensors exist."""
    from research.cifar10.model import ConcatClassifier

    width = 64 if tiny else 768

    def client_data(client_index: int, spec: ExperimentSpec) -> Triple:
        real = _load_partition(Path(spec.data_dir) / "ag_news", client_index)
        if real is not None:
            return real
        gen = torch.Generator().manual_seed(spec.data_seed * 31 + client_index)
        n = spec.samples_per_client
        targets = torch.randint(0, 4, (n,), generator=gen)
        band = (vocab_size - 1) // 4
        topical = 1 + targets.view(-1, 1) * band + torch.randint(0, band, (n, seq_len), generator=gen)
        generic = torch.randint(1, vocab_size, (n, seq_len), generator=gen)
        tokens = torch.where(torch.rand(n, seq_len, generator=gen) < 0.5, topical, generic)
        return split_three_ways(TensorDataset(tokens, targets), spec.data_seed + client_index)

    return Task("ag_news", n_clients, 4, lambda: _BertFeatures(vocab_size, tiny), lambda: nn.Linear(width, 4),
                lambda: ConcatClassifier(width, width, 4), client_data)


# ------------------------------------------------------------------------------------------- FLamby: Fed-Heart-Disease
@task("fed_heart_disease")
def fed_heart_disease(n_clients: int = 4) -> Task:
    """``research/flamby/fed_heart_disease``: 13 tabular features, binary outcome, four hospitals of very different
    sizes; FLamby's baseline is logistic regression.  Reads ``data_dir/fed_heart_disease/client_*`` tensors when
    present (export them from FLamby with ``research.flamby.export``), else draws hospital-shifted synthetic records."""
    from research.cifar10.model import ConcatClassifier

    sizes = (303, 261, 46, 130)  # Cleveland, Hungary, Switzerland, Long Beach

    def client_data(client_index: int, spec: ExperimentSpec) -> Triple:
        real = _load_partition(Path(spec.data_dir) / "fed_heart_disease", client_index)
        if real is not None:
            return real
        gen = torch.Generator().manual_seed(spec.data_seed * 13 + client_index)
        n = max(sizes[client_index % 4], 40)
        x = torch.randn(n, 13, generator=gen) + 0.25 * client_index
        w = torch.randn(13, generator=torch.Generator().manual_seed(5))
        y = ((x - 0.25 * client_index) @ w + 0.5 * torch.randn(n, generator=gen) > 0).long()
        return split_three_ways(TensorDataset(x, y), spec.data_seed + client_index, 0.2, 0.2)

    return Task("fed_heart_disease", n_clients, 2, lambda: nn.Sequential(nn.Linear(13, 16), nn.ReLU()), lambda: nn.Linear(16, 2),
                lambda: ConcatClassifier(16, 16, 2), client_data)
