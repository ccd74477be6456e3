"""Model zoo: example CNNs, CIFAR ResNet-18, BERT encoder."""

from fl4health_b200.models.cnn import ConvNet, MnistNet, MnistNetWithBnAndFrozen, Net
from fl4health_b200.models.bert import BertConfig, BertEncoder, BertForSequenceClassification
from fl4health_b200.models.fused_layers import BatchNormAct2d, Conv2dOverlapWgrad, LinearAct
from fl4health_b200.models.resnet import ResNet18, resnet18_cifar

__all__ = ["BatchNormAct2d", "BertConfig", "BertEncoder", "BertForSequenceClassification", "Conv2dOverlapWgrad", "ConvNet", "LinearAct",
           "MnistNet", "MnistNetWithBnAndFrozen", "Net", "ResNet18", "resnet18_cifar"]
