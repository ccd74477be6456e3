"""Small models used by the example scenarios (MNIST- and CIFAR-shaped inputs)."""

from __future__ import annotations

import torch
from torch import nn

from fl4health_b200.model_bases.parallel_split_models import ParallelFeatureJoinMode, ParallelSplitHeadModule


def _in_channels(dataset: str) -> int:
    return 1 if dataset == "mnist" else 3


class FeatureCnn(nn.Module):
    """conv-pool-conv-pool feature extractor -> flat ``[B, out_dim]`` features (``out_dim`` = 16 * 4 * 4)."""

    out_dim = 256

    def __init__(self, dataset: str = "mnist", batch_norm: bool = False) -> None:
        super().__init__()
        self.conv1 = nn.Conv2d(_in_channels(dataset), 8, 5, padding=2)
        self.bn = nn.BatchNorm2d(8) if batch_norm else nn.Identity()
        self.conv2 = nn.Conv2d(8, 16, 5, padding=2)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = torch.nn.functional.max_pool2d(torch.relu(self.bn(self.conv1(x))), 2)
        x = torch.relu(self.conv2(x))
        return torch.flatten(torch.nn.functional.adaptive_avg_pool2d(x, 4), 1)


class ClassifierHead(nn.Module):
    def __init__(self, in_dim: int = FeatureCnn.out_dim, classes: int = 10) -> None:
        super().__init__()
        self.fc1 = nn.Linear(in_dim, 64)
        self.fc2 = nn.Linear(64, classes)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.fc2(torch.relu(self.fc1(x)))


class SmallCnn(nn.Module):
    """Plain classifier = FeatureCnn + ClassifierHead (state-dict keys ``features.*`` / ``head.*``)."""

    def __init__(self, dataset: str = "mnist", batch_norm: bool = False, frozen_conv: bool = False) -> None:
        super().__init__()
        self.features = FeatureCnn(dataset, batch_norm)
        self.head = ClassifierHead()
        if frozen_conv:  # exercises state_dict != parameters-with-grad (as the reference's SCAFFOLD example model does)
            for p in self.features.conv1.parameters():
                p.requires_grad = False

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.head(self.features(x))


class ConcatHead(ParallelSplitHeadModule):
    """FENDA-style head over concatenated local + global features."""

    def __init__(self, classes: int = 10) -> None:
        super().__init__(ParallelFeatureJoinMode.CONCATENATE)
        self.fc = ClassifierHead(2 * FeatureCnn.out_dim, classes)

    def parallel_output_join(self, local_tensor: torch.Tensor, global_tensor: torch.Tensor) -> torch.Tensor:
        return torch.cat([local_tensor, global_tensor], dim=1)

    def head_forward(self, input_tensor: torch.Tensor) -> torch.Tensor:
        return self.fc(input_tensor)


class Mlp(nn.Module):
    def __init__(self, in_dim: int, hidden: int, out_dim: int) -> None:
        super().__init__()
        self.net = nn.Sequential(nn.Flatten(), nn.Linear(in_dim, hidden), nn.ReLU(), nn.Linear(hidden, out_dim))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.net(x)
