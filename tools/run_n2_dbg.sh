#!/usr/bin/env bash
out=gpurun_out/n2b; mkdir -p $out
export FL4H_LOG_LEVEL=ERROR
tr() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "$@"; }
for coll in auto nccl; do
  FL4H_COLLECTIVES=$coll tr -m examples.run scaffold_example --spmd --rounds 3 2>&1 | grep '^{"scenario"' | cut -c1-220 | sed "s/^/scaffold $coll /"
done
FL4H_COLLECTIVES=auto FL4H_SCAFFOLD_DEBUG=1 tr tools/dbg_scaffold_spmd.py 2>&1 | grep "DBG" | head -40
tr -m examples.run fedpm_example --spmd --rounds 3 2>&1 | grep '^{"scenario"' | cut -c1-220 | sed "s/^/fedpm /"
