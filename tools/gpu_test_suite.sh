#!/bin/bash
# Full GPU suite (kernel numerics, engine, method scenarios with CUDA graphs).
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | cut -c1-250 | tee gpurun_out/gpu_tests_final.log
