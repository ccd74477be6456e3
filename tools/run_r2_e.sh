#!/usr/bin/env bash
out=gpurun_out/r2e; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_dp.py tests/test_gpu_layer_norm.py -m gpu -q --tb=short > $out/pytest_new.log 2>&1; echo "== new tests rc=$?"; tail -25 $out/pytest_new.log | cut -c1-260
for kind in basic fedprox; do
  KINETO_GC=1 KINETO_DTYPE=fp32 KINETO_CLIENT=$kind timeout 300 python benchmarks/kineto_step.py > $out/kineto_$kind.txt 2>&1; echo "== kineto $kind rc=$?"; grep -A14 "tensors with autograd" $out/kineto_$kind.txt | cut -c1-250; grep "replays=" $out/kineto_$kind.txt
done
FL4H_TRACE=1 timeout 600 python bench.py --config fedper_ditto_dp --steps 20 --warmup 5 --skip-e2e --skip-extra-dtype > $out/trace_dp.json 2> $out/trace_dp.err; echo "== dp cfg rc=$?"; grep '^{"dtype"' $out/trace_dp.err | cut -c1-600; grep -B2 -A12 "Traceback" $out/trace_dp.err | head -60 | cut -c1-250
timeout 420 python bench.py --config bert_fedadam --steps 5 --warmup 3 > $out/bench_bert.json 2> $out/bench_bert.err; echo "== bert rc=$?"; tail -1 $out/bench_bert.json | cut -c1-700; tail -3 $out/bench_bert.err | cut -c1-300
