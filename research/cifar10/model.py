"""CIFAR-10 pFL benchmark networks: the 2-conv / 2-fc ``ConvNet`` (8.47 M parameters at the default width) as a
feature extractor + classifier pair, so one definition serves the plain, sequentially split (FedPer / MOON / FedRep)
and parallel (FENDA / PerFCL / APFL) model families (parity of architecture: ``research/cifar10/model.py:10-166``)."""

from __future__ import annotations

import torch
from torch import nn

from fl4health_b200.model_bases.parallel_split_models import ParallelFeatureJoinMode, ParallelSplitHeadModule


class ConvFeatures(nn.Module):
    """conv5-BN-ReLU-pool ×2 -> flat ``64 * (h/4) * (w/4)`` features."""

    def __init__(self, in_channels: int = 3, use_bn: bool = True) -> None:
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, 32, 5, padding=2)
        self.bn1 = nn.BatchNorm2d(32) if use_bn else nn.Identity()
        self.conv2 = nn.Conv2d(32, 64, 5, padding=2)
        self.bn2 = nn.BatchNorm2d(64) if use_bn else nn.Identity()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = torch.max_pool2d(torch.relu(self.bn1(self.conv1(x))), 2)
        x = torch.max_pool2d(torch.relu(self.bn2(self.conv2(x))), 2)
        return torch.flatten(x, 1)


def feature_dim(h: int = 32, w: int = 32) -> int:
    return 64 * (h // 4) * (w // 4)


class MlpClassifier(nn.Module):
    def __init__(self, in_dim: int, hidden: int = 2048, class_num: int = 10, dropout: float = 0.0) -> None:
        super().__init__()
        self.fc1 = nn.Linear(in_dim, hidden)
        self.fc2 = nn.Linear(hidden, class_num)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.fc2(self.dropout(torch.relu(self.fc1(self.dropout(x)))))


class ConcatClassifier(ParallelSplitHeadModule):
    """Head over concatenated local + global features (twice the input width of ``MlpClassifier``)."""

    def __init__(self, in_dim: int, hidden: int = 2048, class_num: int = 10, dropout: float = 0.0) -> None:
        super().__init__(ParallelFeatureJoinMode.CONCATENATE)
        self.classifier = MlpClassifier(2 * in_dim, hidden, class_num, dropout)

    def parallel_output_join(self, local_tensor: torch.Tensor, global_tensor: torch.Tensor) -> torch.Tensor:
        return torch.cat((local_tensor, global_tensor), dim=1)

    def head_forward(self, input_tensor: torch.Tensor) -> torch.Tensor:
        return self.classifier(input_tensor)


class SplitClassifier(nn.Module):
    """``head(features(x))`` with stable attribute names: ``features`` is the layer MMD penalties hook, and the
    ``features.*`` / ``head.*`` state-dict prefixes are what warm-start mappings refer to."""

    def __init__(self, features: nn.Module, head: nn.Module) -> None:
        super().__init__()
        self.features = features
        self.head = head

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.head(self.features(x))


def conv_net(in_channels: int = 3, h: int = 32, w: int = 32, hidden: int = 2048, class_num: int = 10, use_bn: bool = True,
             dropout: float = 0.0) -> SplitClassifier:
    return SplitClassifier(ConvFeatures(in_channels, use_bn), MlpClassifier(feature_dim(h, w), hidden, class_num, dropout))
