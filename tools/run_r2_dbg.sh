#!/usr/bin/env bash
out=gpurun_out/r2h; mkdir -p $out
export FL4H_LOG_LEVEL=ERROR
one() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "$@"; }
for sc in scaffold_example fedpm_example basic_example fedprox_example; do
  timeout 200 python -m examples.run $sc --rounds 3 --clients 2 --device cuda 2>&1 | grep '^{"scenario"' | cut -c1-200 | sed "s/^/sim  /"
  one -m examples.run $sc --spmd --rounds 3 2>&1 | grep '^{"scenario"' | cut -c1-200 | sed "s/^/spmd1 /"
done
FL4H_FUSED_OPT=0 one -m examples.run scaffold_example --spmd --rounds 3 2>&1 | grep '^{"scenario"' | cut -c1-200 | sed "s/^/spmd1 nofusedopt /"
FL4H_ARENA=0 one -m examples.run scaffold_example --spmd --rounds 3 2>&1 | grep '^{"scenario"' | cut -c1-200 | sed "s/^/spmd1 noarena /"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
