#!/usr/bin/env bash
out=gpurun_out/r2g; mkdir -p $out
timeout 300 python benchmarks/bert_block_bench.py > $out/bert_block_bench.txt 2>&1; echo "== block bench rc=$?"; grep BLOCK $out/bert_block_bench.txt | cut -c1-400 || tail -5 $out/bert_block_bench.txt
timeout 420 python bench.py --config bert_fedadam --steps 10 --warmup 3 > $out/bench_bert.json 2> $out/bench_bert.err; echo "== bert rc=$?"; tail -1 $out/bench_bert.json | cut -c1-900; tail -3 $out/bench_bert.err | cut -c1-300
FL4H_TC_ATTENTION=0 FL4H_LN_KERNEL=0 timeout 420 python bench.py --config bert_fedadam --steps 10 --warmup 3 > $out/bench_bert_library_blocks.json 2> $out/bench_bert_lib.err; echo "== bert (library attention + LN) rc=$?"; tail -1 $out/bench_bert_library_blocks.json | cut -c1-400
# ncu: one capture per new kernel family (single GPU, kernels selected by name, few launches)
for k in attention_fwd_kernel attention_bwd_kernel ln_fwd_kernel ln_bwd_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -c 2 -o $out/ncu_$k -f python benchmarks/bert_block_bench.py > $out/ncu_$k.log 2>&1; echo "== ncu $k rc=$?"
  ncu -i $out/ncu_$k.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_pick.py > $out/ncu_$k.summary.csv 2>/dev/null; head -3 $out/ncu_$k.summary.csv | cut -c1-300
done
