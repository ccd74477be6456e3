#!/bin/bash
# Phase timelines of the fused BN kernels at three ResNet-18 / CIFAR layer shapes.
for shape in "32 64 32 32" "32 128 16 16" "32 512 4 4"; do timeout 100 python benchmarks/bn_phases.py $shape 2>&1 | tail -6; done
