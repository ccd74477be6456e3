#!/usr/bin/env bash
# One 8-GPU session: collectives correctness + timing (both data paths), headline bench, reference arm.
mkdir -p gpurun_out/n8
NP=${NP:-8}
tr() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "$@"; }
for nv in 1 0; do
  FL4H_NVLS=$nv tr tests/fused_worker.py gpurun_out/n8/fused_nvls$nv.json 11173888 > gpurun_out/n8/fused_nvls$nv.log 2>&1
  tail -1 gpurun_out/n8/fused_nvls$nv.log | cut -c1-1500
  FL4H_NVLS=$nv tr benchmarks/fused_tune.py 2>&1 | grep TUNE | cut -c1-900
done
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT tr bench.py --gpus $NP --steps 20 --warmup 5 > gpurun_out/n8/native.json 2> gpurun_out/n8/native.err
grep -m2 -i "nvls" gpurun_out/n8/native.err | cut -c1-200
tail -c 1500 gpurun_out/n8/native.json
tr bench.py --impl reference --gpus $NP --steps 10 --warmup 3 > gpurun_out/n8/ref.json 2> gpurun_out/n8/ref.err
tail -c 600 gpurun_out/n8/ref.json
