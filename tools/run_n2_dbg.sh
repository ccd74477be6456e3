#!/usr/bin/env bash
export FL4H_LOG_LEVEL=ERROR
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tools/dbg_scaffold_spmd.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -40 | cut -c1-260
