mkdir -p gpurun_out/r2c
run() { timeout 200 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) benchmarks/fused_tune.py 2>&1 | grep TUNE; }
NP=${NP:-2}
run FL4H_NVLS=0
run FL4H_NVLS=1 FL4H_NVLS_UNROLL=4
run FL4H_NVLS=1 FL4H_NVLS_UNROLL=2
run FL4H_NVLS=1 FL4H_NVLS_UNROLL=4
run FL4H_NVLS=1 FL4H_NVLS_UNROLL=8
run FL4H_NVLS=1 FL4H_NVLS_UNROLL=8 FL4H_COLL_GRID=64
run FL4H_NVLS=1 FL4H_NVLS_UNROLL=4 FL4H_COLL_GRID=32
