"""Model zoo: example CNNs, CIFAR ResNet-18, BERT encoder."""

from fl4health_b200.models.cnn import ConvNet, MnistNet, MnistNetWithBnAndFrozen, Net
from fl4health_b200.models.resnet import ResNet18, resnet18_cifar

__all__ = ["ConvNet", "MnistNet", "MnistNetWithBnAndFrozen", "Net", "ResNet18", "resnet18_cifar"]
