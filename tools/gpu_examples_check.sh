#!/bin/bash
# Method clients through the engine with CUDA graphs on a GPU box (each scenario in its own process group so a device-side
# assert in one cannot poison the others), plus the optimizer inf-parameter regression test.
timeout 120 python -m pytest tests/test_gpu_flat_ops.py -m gpu -q -k "infinite or sgd_step or adamw or mt_" 2>&1 | tail -2 | cut -c1-200
timeout 400 python -m pytest tests/test_gpu_examples.py -m gpu -q 2>&1 | tail -6 | cut -c1-250
