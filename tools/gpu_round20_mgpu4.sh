#!/bin/bash
mkdir -p gpurun_out
SUM='import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ["n_gpus","value","ms_per_step","stages_ms","exact_fallback_rows"]}, d["e2e"], d["roofline"]["frac"])'
echo "=== mgpu_check world 4"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py 2>&1 | grep -E "MGPU_|Error|error" | head
echo "=== bench --gpus 4"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 10 --warmup 3 2>&1 | grep '^{' | tail -1 | tee gpurun_out/r01o_bench_n4.json | python -c "$SUM"
df -h /dev/shm | tail -1
