#!/usr/bin/env bash
# Final 8-GPU session of round 2: headline (fused / NCCL A/B), reference arm, SCAFFOLD+FedProx config, BERT FedAdam, and the
# SPMD examples that exercise stock BatchNorm + packed payloads under fused collectives.
out=gpurun_out/n8f; mkdir -p $out
NP=${NP:-8}
tr() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "$@"; }
tr bench.py --gpus $NP --steps 100 --warmup 5 > $out/native.json 2> $out/native.err; echo "== native rc=$?"; tail -1 $out/native.json | cut -c1-400
tr bench.py --gpus $NP --steps 60 --warmup 5 --collectives nccl --skip-e2e --skip-extra-dtype > $out/native_nccl.json 2> $out/native_nccl.err; echo "== native nccl rc=$?"; tail -1 $out/native_nccl.json | cut -c1-300
tr bench.py --impl reference --gpus $NP --steps 8 --warmup 3 > $out/ref.json 2> $out/ref.err; echo "== reference rc=$?"; tail -1 $out/ref.json | cut -c1-500
tr bench.py --gpus $NP --config scaffold_fedprox --steps 30 --warmup 5 --skip-e2e > $out/scaffold_fedprox.json 2> $out/scaffold_fedprox.err; echo "== scaffold_fedprox rc=$?"
tr bench.py --gpus $NP --config bert_fedadam --steps 10 --warmup 3 > $out/bert.json 2> $out/bert.err; echo "== bert rc=$?"; tail -1 $out/bert.json | cut -c1-500
python - <<'PY'
import json
for f in ("native", "native_nccl", "scaffold_fedprox"):
    try:
        d = json.loads(open(f"gpurun_out/n8f/{f}.json").read().strip().splitlines()[-1])
        print(" ", f, d["config"].get("collectives"), "ms/round", round(d["ms_per_step"], 3), "value", round(d["value"], 1), "e2e", (d.get("e2e") or {}).get("value"),
              "bf16", (d.get("bf16") or {}).get("value"), {k: round(v["ms_per_step"], 3) for k, v in d.get("variants", {}).items()})
    except Exception as e:
        print(" ", f, "parse failed", e)
PY
export FL4H_LOG_LEVEL=ERROR
for sc in scaffold_example fedpm_example; do
  tr -m examples.run $sc --spmd --rounds 3 2>&1 | grep '^{"scenario"' | cut -c1-260
done
