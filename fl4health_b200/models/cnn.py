"""Small CNNs used by the examples (architectures & state-dict key names match the reference's example models so
checkpoints are interchangeable: ``examples/models/cnn_model.py:6-62``, ``research/cifar10/model.py:10-47``)."""

from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn


class Net(nn.Module):
    """LeNet-style CIFAR-10 network (62 006 parameters): 2x(conv5x5 + pool) -> 120 -> 84 -> 10."""

    def __init__(self, in_channels: int = 3, num_classes: int = 10) -> None:
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, 6, kernel_size=5)
        self.pool = nn.MaxPool2d(kernel_size=2, stride=2)
        self.conv2 = nn.Conv2d(6, 16, kernel_size=5)
        self.fc1 = nn.Linear(16 * 5 * 5, 120)
        self.fc2 = nn.Linear(120, 84)
        self.fc3 = nn.Linear(84, num_classes)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.pool(F.relu(self.conv1(x)))
        x = self.pool(F.relu(self.conv2(x)))
        x = torch.flatten(x, 1)
        return self.fc3(F.relu(self.fc2(F.relu(self.fc1(x)))))


class MnistNet(nn.Module):
    """MNIST CNN (~35k parameters).  The final ReLU on the logits is intentional (matches the reference)."""

    def __init__(self) -> None:
        super().__init__()
        self.conv1 = nn.Conv2d(1, 8, kernel_size=5)
        self.pool = nn.MaxPool2d(kernel_size=2, stride=2)
        self.conv2 = nn.Conv2d(8, 16, kernel_size=5)
        self.fc1 = nn.Linear(16 * 4 * 4, 120)
        self.fc2 = nn.Linear(120, 10)

    def features(self, x: torch.Tensor) -> torch.Tensor:
        x = self.pool(F.relu(self.conv1(x)))
        return self.pool(F.relu(self.conv2(x)))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = torch.flatten(self.features(x), 1)
        return F.relu(self.fc2(F.relu(self.fc1(x))))


class MnistNetWithBnAndFrozen(MnistNet):
    """MNIST CNN + BatchNorm + (optionally) a frozen first conv: exercises ``state_dict != parameters`` paths
    (SCAFFOLD variates cover trainable params only, exchange covers all state)."""

    def __init__(self, freeze_cnn_layer: bool = True) -> None:
        super().__init__()
        self.bn = nn.BatchNorm2d(num_features=16)
        if freeze_cnn_layer:
            self.conv1.requires_grad_(False)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = torch.flatten(self.bn(self.features(x)), 1)
        return F.relu(self.fc2(F.relu(self.fc1(x))))


class ConvNet(nn.Module):
    """Research CIFAR-10 network: 2x(conv5x5 + BN + ReLU + pool) -> fc(hidden) -> classes (8.47 M params at 32x32)."""

    def __init__(
        self, in_channels: int, h: int = 32, w: int = 32, hidden: int = 2048, class_num: int = 10,
        use_bn: bool = True, dropout: float = 0.0,
    ) -> None:
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, 32, 5, padding=2)
        self.conv2 = nn.Conv2d(32, 64, 5, padding=2)
        self.use_bn = use_bn
        if use_bn:
            self.bn1 = nn.BatchNorm2d(32)
            self.bn2 = nn.BatchNorm2d(64)
        self.fc1 = nn.Linear((h // 4) * (w // 4) * 64, hidden)
        self.fc2 = nn.Linear(hidden, class_num)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(2)
        self.dropout_layer = nn.Dropout(p=dropout)
        self.flatten = nn.Flatten()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.conv1(x)
        x = self.maxpool(self.relu(self.bn1(x) if self.use_bn else x))
        x = self.conv2(x)
        x = self.maxpool(self.relu(self.bn2(x) if self.use_bn else x))
        x = self.dropout_layer(self.flatten(x))
        x = self.dropout_layer(self.relu(self.fc1(x)))
        return self.fc2(x)
