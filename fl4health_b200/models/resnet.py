"""ResNet-18 for 32x32 inputs (the CIFAR variant: 3x3 stem, no max-pool) — 11.17 M parameters.

The reference only ever uses ``torchvision.models.resnet18`` (``research/rxrx1/utils.py:91``); ``BASELINE.json``'s
headline config is "CIFAR-10 ResNet-18", so the framework ships its own.  State-dict key names follow torchvision
(``conv1/bn1/layer{1-4}.{0,1}.{conv,bn}{1,2}/downsample.{0,1}/fc``) so torchvision weights load with
``strict=True`` when ``imagenet_stem=True``.
"""

from __future__ import annotations

import torch
from torch import nn

from fl4health_b200.engine import streams
from fl4health_b200.models.fused_layers import BatchNormAct2d, Conv2dOverlapWgrad, TcConv2d, bn_act, conv_bn_act


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, in_planes: int, planes: int, stride: int = 1) -> None:
        super().__init__()
        # tcgen05 implicit-GEMM convolutions whose epilogue reduces the BatchNorm statistics (ops/csrc/conv_tc.cu)
        self.conv1 = TcConv2d(in_planes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn1 = BatchNormAct2d(planes, relu=True)  # bn + relu fused
        self.conv2 = TcConv2d(planes, planes, 3, stride=1, padding=1, bias=False)
        self.bn2 = BatchNormAct2d(planes, relu=True)  # bn + residual add + relu fused
        self.downsample: nn.Module | None = None
        if stride != 1 or in_planes != planes:
            self.downsample = nn.Sequential(
                TcConv2d(in_planes, planes, 1, stride=stride, bias=False), BatchNormAct2d(planes, relu=False)
            )

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.downsample is None:
            out = conv_bn_act(self.conv1, self.bn1, x)
            return conv_bn_act(self.conv2, self.bn2, out, residual=x)
        if x.is_cuda and streams.overlap_enabled():
            # the projection shortcut is independent of conv1/bn1/conv2: run it on a side stream (its backward follows it
            # there), joining right before the residual add; both branches are small enough to share the 148 SMs
            main = torch.cuda.current_stream(x.device)
            side = streams.branch_stream(x.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                identity = conv_bn_act(self.downsample[0], self.downsample[1], x, relu=False)
            x.record_stream(side)
            out = conv_bn_act(self.conv1, self.bn1, x)
            main.wait_stream(side)
            identity.record_stream(main)
            return conv_bn_act(self.conv2, self.bn2, out, residual=identity)
        identity = conv_bn_act(self.downsample[0], self.downsample[1], x, relu=False)
        out = conv_bn_act(self.conv1, self.bn1, x)
        return conv_bn_act(self.conv2, self.bn2, out, residual=identity)


class ResNet18(nn.Module):
    def __init__(self, num_classes: int = 10, in_channels: int = 3, imagenet_stem: bool = False) -> None:
        super().__init__()
        if imagenet_stem:
            self.conv1 = Conv2dOverlapWgrad(in_channels, 64, 7, stride=2, padding=3, bias=False)
            self.maxpool: nn.Module = nn.MaxPool2d(3, stride=2, padding=1)
        else:
            self.conv1 = TcConv2d(in_channels, 64, 3, stride=1, padding=1, bias=False)  # Cin = 3: CUDA-core stem kernels
            self.maxpool = nn.Identity()
        self.bn1 = BatchNormAct2d(64, relu=True)
        widths, strides = (64, 128, 256, 512), (1, 2, 2, 2)
        in_planes = 64
        for idx, (planes, stride) in enumerate(zip(widths, strides), start=1):
            setattr(self, f"layer{idx}", nn.Sequential(BasicBlock(in_planes, planes, stride), BasicBlock(planes, planes)))
            in_planes = planes
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(512, num_classes)
        for module in self.modules():
            if isinstance(module, nn.Conv2d):
                nn.init.kaiming_normal_(module.weight, mode="fan_out", nonlinearity="relu")

    def forward_features(self, x: torch.Tensor) -> torch.Tensor:
        x = self.maxpool(conv_bn_act(self.conv1, self.bn1, x))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return torch.flatten(self.avgpool(x), 1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.fc(self.forward_features(x))


def resnet18_cifar(num_classes: int = 10) -> ResNet18:
    return ResNet18(num_classes=num_classes)
