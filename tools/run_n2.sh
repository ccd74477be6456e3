#!/usr/bin/env bash
# One 2-GPU session: multi-GPU tests, headline + variant benches (fused vs NCCL A/B), clients spanning 2 GPUs, FedPM bit vote.
out=gpurun_out/n2; mkdir -p $out
NP=${NP:-2}
tr() { timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "$@"; }
timeout 300 python benchmarks/bert_block_bench.py > $out/bert_block_bench.txt 2>&1; grep BLOCK $out/bert_block_bench.txt | cut -c1-400
timeout 200 ncu --set full --clock-control none --import-source on -k regex:ln_bwd_kernel -c 2 -o $out/ncu_ln_bwd_kernel -f python benchmarks/bert_block_bench.py > $out/ncu_ln_bwd.log 2>&1
ncu -i $out/ncu_ln_bwd_kernel.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_pick.py > $out/ncu_ln_bwd_kernel.summary.csv 2>/dev/null
timeout 600 python -m pytest tests/test_multigpu_fused.py -m gpu -q --tb=short > $out/pytest_multigpu.log 2>&1; echo "== multigpu tests rc=$?"; tail -3 $out/pytest_multigpu.log | cut -c1-200
tr bench.py --gpus $NP --steps 50 --warmup 5 > $out/bench_fused.json 2> $out/bench_fused.err; echo "== headline fused rc=$?"; tail -1 $out/bench_fused.json | cut -c1-500
tr bench.py --gpus $NP --steps 50 --warmup 5 --collectives nccl --skip-extra-dtype > $out/bench_nccl.json 2> $out/bench_nccl.err; echo "== headline nccl rc=$?"; tail -1 $out/bench_nccl.json | cut -c1-300
for cfg in scaffold_fedprox fedper_ditto_dp; do
  for coll in auto nccl; do
    tr bench.py --gpus $NP --config $cfg --steps 20 --warmup 5 --skip-e2e --collectives $coll > $out/bench_${cfg}_$coll.json 2> $out/bench_${cfg}_$coll.err; echo "== $cfg $coll rc=$?"
    python - <<PY
import json
try:
    d=json.loads(open("$out/bench_${cfg}_$coll.json").read().strip().splitlines()[-1])
    print("  ->", d["config"]["collectives"], round(d["ms_per_step"],3), {k: round(v["ms_per_step"],3) for k,v in d.get("variants",{}).items()})
except Exception as e: print("  parse failed", e)
PY
  done
done
tr bench.py --gpus $NP --config bert_fedadam --steps 10 --warmup 3 > $out/bench_bert.json 2> $out/bench_bert.err; echo "== bert rc=$?"; tail -1 $out/bench_bert.json | cut -c1-600
# clients spanning 2 GPUs (NCCL gradient all-reduce in the optimizer pre-hook) and the FedPM packed-bit vote, on real NVLink
tr -m examples.run ditto_example --spmd --ranks-per-client 2 --rounds 3 > $out/ditto_rpc2.log 2>&1; echo "== ditto ranks-per-client 2 rc=$?"; grep '^{"scenario"' $out/ditto_rpc2.log | cut -c1-300
tr -m examples.run fedllm_example --spmd --ranks-per-client 2 --rounds 2 > $out/fedllm_rpc2.log 2>&1; echo "== fedllm ranks-per-client 2 rc=$?"; grep '^{"scenario"' $out/fedllm_rpc2.log | cut -c1-300
tr -m examples.run fedpm_example --spmd --rounds 3 > $out/fedpm.log 2>&1; echo "== fedpm spmd rc=$?"; grep '^{"scenario"' $out/fedpm.log | cut -c1-300
tr -m examples.run scaffold_example --spmd --rounds 3 > $out/scaffold.log 2>&1; echo "== scaffold spmd rc=$?"; grep '^{"scenario"' $out/scaffold.log | cut -c1-300
FL4H_TEST_VARIANT=ditto tr tests/client_group_worker.py $out/cg_ditto 2 > $out/cg_ditto.log 2>&1; echo "== client_group worker (ditto, NCCL) rc=$?"
python - <<'PY'
import json
try:
    a, b = (json.load(open(f"gpurun_out/n2/cg_ditto.rank{r}")) for r in (0, 1))
    print("  replicas agree before aggregate:", {k: (a["pre_aggregate"][k], b["pre_aggregate"][k]) for k in a["pre_aggregate"]})
except Exception as e: print("  no result", e)
PY
