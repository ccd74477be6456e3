#!/usr/bin/env bash
# Phase breakdown (FL4H_TRACE=1: device ms per round phase on stderr) of the bench variants + the re-worked GPU tests.
out=gpurun_out/r2c; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_flat_ops.py tests/test_gpu_conv.py -m gpu -q --tb=short -k "fused_bn_matches or epilogue_statistics or packed_vote" > $out/pytest_subset.log 2>&1; echo "== subset rc=$?"; tail -5 $out/pytest_subset.log | cut -c1-300
for cfg in cifar_fedavg scaffold_fedprox fedper_ditto_dp; do
  FL4H_TRACE=1 timeout 420 python bench.py --config $cfg --steps 10 --warmup 3 --skip-e2e --skip-extra-dtype > $out/trace_$cfg.json 2> $out/trace_$cfg.err; echo "== $cfg rc=$?"
  grep '^{"dtype"' $out/trace_$cfg.err | cut -c1-1200
done
for kind in basic fedprox; do
  KINETO_DTYPE=fp32 KINETO_CLIENT=$kind timeout 300 python benchmarks/kineto_step.py > $out/kineto_$kind.txt 2>&1; echo "== kineto $kind rc=$?"; head -3 $out/kineto_$kind.txt | cut -c1-250
done
