#!/bin/bash
mkdir -p gpurun_out
SUM='import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ["value","ms_per_step","stages_ms","exact_fallback_rows"]}, d["e2e"]["ms_per_step"], d["roofline"]["frac"])'
echo "=== newref gpu tests"; timeout 280 python -m pytest tests/test_newref_gpu.py -q -x --tb=short 2>&1 | tail -25
for K in 5 6 4; do
echo "=== bench config3 kernel=$K"; timeout 120 python bench.py --kernel $K --steps 3 --warmup 2 --no-cpu-baseline --no-predict 2>&1 | tail -1 | python -c "$SUM"
done
echo "=== bench config2 kernel=5"; timeout 120 python bench.py --workload config2 --kernel 5 --steps 3 --warmup 2 --no-cpu-baseline --no-predict 2>&1 | tail -1 | python -c "$SUM"
