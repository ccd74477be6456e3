#!/usr/bin/env bash
out=gpurun_out/r2d; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q --tb=short -x > $out/pytest_dp.log 2>&1; echo "== dp tests rc=$?"; tail -12 $out/pytest_dp.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > $out/pytest_gpu.log 2>&1; echo "== gpu suite rc=$?"; tail -4 $out/pytest_gpu.log | cut -c1-300
for cfg in cifar_fedavg scaffold_fedprox fedper_ditto_dp; do
  FL4H_TRACE=1 timeout 600 python bench.py --config $cfg --steps 20 --warmup 5 --skip-e2e --skip-extra-dtype > $out/trace_$cfg.json 2> $out/trace_$cfg.err; echo "== $cfg rc=$?"
  grep '^{"dtype"' $out/trace_$cfg.err | cut -c1-1200
  python - <<PY
import json
d=json.loads(open("$out/trace_$cfg.json").read().strip().splitlines()[-1])
print("  ->", d["metric"], round(d["ms_per_step"],3), {k: round(v["ms_per_step"],3) for k,v in d.get("variants",{}).items()})
PY
done
grep -c "AccumulateGrad" $out/trace_scaffold_fedprox.err
