#!/usr/bin/env bash
# Each group in its own process under its own timeout: a hung kernel cannot take the whole box budget with it.
mkdir -p gpurun_out/conv
for sel in "forward and tf32" "forward and bf16" "backward and tf32" "backward and bf16" "autograd"; do
  name=$(echo "$sel" | tr ' ' '_')
  timeout 240 python -m pytest tests/test_gpu_conv.py -q --tb=line -k "$sel" > gpurun_out/conv/$name.log 2>&1
  echo "== $sel rc=$?"; tail -6 gpurun_out/conv/$name.log | cut -c1-220
done
timeout 300 python benchmarks/conv_bench.py > gpurun_out/conv/bench.txt 2>&1; grep CONV gpurun_out/conv/bench.txt || tail -5 gpurun_out/conv/bench.txt
for sel in "stem" "resnet18_step"; do
  timeout 240 python -m pytest tests/test_gpu_conv.py -q --tb=short -k "$sel" > gpurun_out/conv/$sel.log 2>&1
  echo "== $sel rc=$?"; tail -8 gpurun_out/conv/$sel.log | cut -c1-220
done
