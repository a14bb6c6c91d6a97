#!/bin/bash
# multi-GPU: sharded get_reference == single GPU, bench at 2 and 4 ranks
mkdir -p gpurun_out
TAG=${1:-r01l}
N=${2:-4}
SUM='import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ["n_gpus","value","ms_per_step","stages_ms","exact_fallback_rows"]}, d["e2e"]["ms_per_step"], d["roofline"]["frac"])'
echo "=== mgpu_check world $N"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py 2>&1 | grep -E "MGPU_|Error|error" | head
for W in 2 $N; do
echo "=== bench --gpus $W"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $W --steps 5 --warmup 3 2>&1 | grep '^{' | tail -1 | tee gpurun_out/${TAG}_bench_n$W.json | python -c "$SUM"
done
echo "=== reference arm under torchrun"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>&1 | grep '^{' | tail -1 | cut -c1-300
