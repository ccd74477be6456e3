#!/usr/bin/env bash
# Round-2 single-GPU sweep: GPU test-suite, the BASELINE configurations through bench.py, conv micro-benchmarks.
# Every step runs under its own timeout and logs into gpurun_out/r2b/.
out=gpurun_out/r2b; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q --tb=short > $out/pytest_gpu.log 2>&1; echo "== pytest gpu rc=$?"; tail -4 $out/pytest_gpu.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 > $out/bench_headline.json 2> $out/bench_headline.err; echo "== headline rc=$?"; tail -1 $out/bench_headline.json | cut -c1-600
for cfg in scaffold_fedprox fedper_ditto_dp; do
  timeout 420 python bench.py --config $cfg --steps 10 --warmup 3 > $out/bench_$cfg.json 2> $out/bench_$cfg.err; echo "== $cfg rc=$?"
  tail -1 $out/bench_$cfg.json | cut -c1-900; tail -3 $out/bench_$cfg.err | cut -c1-300
done
timeout 420 python bench.py --config bert_fedadam --steps 5 --warmup 3 > $out/bench_bert.json 2> $out/bench_bert.err; echo "== bert rc=$?"; tail -1 $out/bench_bert.json | cut -c1-900; tail -3 $out/bench_bert.err | cut -c1-300
timeout 300 python benchmarks/conv_bench.py > $out/conv_bench.txt 2>&1; grep CONV $out/conv_bench.txt | cut -c1-200 || tail -5 $out/conv_bench.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $out/smi.csv 2>&1
