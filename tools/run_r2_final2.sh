#!/usr/bin/env bash
out=gpurun_out/r2j; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > $out/pytest_gpu.log 2>&1; echo "== gpu suite rc=$?"; tail -3 $out/pytest_gpu.log | cut -c1-200
timeout 300 python benchmarks/convergence_parity.py --rounds 30 --lr 0.003 > $out/parity.txt 2>&1; echo "== parity rc=$?"; grep PARITY $out/parity.txt | cut -c1-420 || tail -5 $out/parity.txt
FL4H_TRACE=1 timeout 420 python bench.py --config fedper_ditto_dp --steps 20 --warmup 5 --skip-e2e > $out/fedper.json 2> $out/fedper.err; echo "== fedper cfg rc=$?"; grep '^{"dtype"' $out/fedper.err | head -1 | cut -c1-400
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2j/fedper.json").read().strip().splitlines()[-1]); print("  ->", round(d["ms_per_step"],3), {k: round(v["ms_per_step"],3) for k,v in d.get("variants",{}).items()})
PY
timeout 600 python bench.py > $out/bench_default.json 2> $out/bench_default.err; echo "== bench default rc=$?"; tail -1 $out/bench_default.json | cut -c1-1500
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
