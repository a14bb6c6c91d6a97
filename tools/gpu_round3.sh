#!/bin/bash
mkdir -p gpurun_out
SUM='import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ["value","ms_per_step","stages_ms","exact_fallback_rows"]}, d["e2e"]["ms_per_step"])'
echo "=== newref gpu tests"; timeout 120 python -m pytest tests/test_newref_gpu.py -x -q --tb=short 2>&1 | tail -15
for CFG in 4,2 8,2 4,3; do
echo "=== bench config3 bulk rerank warps,stages=$CFG"; WCX_RERANK_BULK=$CFG timeout 120 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-predict 2>&1 | tail -1 | python -c "$SUM"
done
echo "=== bench config2"; timeout 120 python bench.py --workload config2 --steps 3 --warmup 2 --no-cpu-baseline --no-predict 2>&1 | tail -1 | python -c "$SUM"
