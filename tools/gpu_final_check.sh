#!/bin/bash
# End-of-change check on a GPU box: full GPU suite, headline bench, smoke().
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | cut -c1-300 | tee gpurun_out/gpu_tests_final.log
python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_1gpu_final.json; cut -c1-330 gpurun_out/bench_1gpu_final.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
