#!/usr/bin/env bash
out=gpurun_out/r2f; mkdir -p $out
timeout 180 python -m pytest tests/test_gpu_attention.py -m gpu -q --tb=short -x > $out/pytest_att.log 2>&1; echo "== attention tests rc=$?"; tail -30 $out/pytest_att.log | cut -c1-260
