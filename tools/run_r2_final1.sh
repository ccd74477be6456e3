#!/usr/bin/env bash
out=gpurun_out/r2i; mkdir -p $out
timeout 300 python benchmarks/convergence_parity.py --rounds 30 > $out/parity.txt 2>&1; echo "== parity rc=$?"; grep PARITY $out/parity.txt | cut -c1-420 || tail -5 $out/parity.txt
OUT_DIR=$out PER_TOOL_TIMEOUT=400 bash tools/sanitize_r2.sh > $out/sanitize.log 2>&1; echo "== sanitizers rc=$?"; cat $out/sanitize.log | cut -c1-200 | tail -40
