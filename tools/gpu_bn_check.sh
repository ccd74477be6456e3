#!/bin/bash
# BN kernel loop: numerics tests (everything but tcgen05), the forward phase timeline and the captured step time.
timeout 250 python -m pytest tests/test_gpu_flat_ops.py tests/test_gpu_engine.py -m gpu -x -q -k "not tcgen05" 2>&1 | tail -2 | cut -c1-200
timeout 100 python benchmarks/bn_phases.py 2>&1 | tail -4
python benchmarks/profile_step.py 2>&1 | grep -i 'train_step_ms' | cut -c1-200
