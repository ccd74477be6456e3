#!/usr/bin/env bash
export FL4H_LOG_LEVEL=DEBUG
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tools/dbg_scaffold_spmd.py > gpurun_out/n2b_scaffold_full.log 2>&1
grep -n -i "error\|fail\|exception\|traceback\|DBG\|assert" gpurun_out/n2b_scaffold_full.log | head -40 | cut -c1-300
