#!/bin/bash
# round-14: null ratios of finished row blocks on a high-priority side stream next to the re-rank
mkdir -p gpurun_out
TAG=${1:-r01j}
SUM='import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ["value","ms_per_step","stages_ms","exact_fallback_rows"]}, d["e2e"]["ms_per_step"], d["roofline"]["frac"])'
echo "=== gpu tests (newref)"; timeout 900 python -m pytest tests/test_newref_gpu.py -q -x --tb=short 2>&1 | tail -15
echo "=== bench config3 side stream"; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-predict 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench.json | python -c "$SUM"
echo "=== bench config3 serial"; WCX_SERIAL_NULLS=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-predict 2>&1 | tail -1 | python -c "$SUM"
echo "=== bench config3 fused"; WCX_FUSED_NULLS=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-predict 2>&1 | tail -1 | python -c "$SUM"
echo "=== bench config2"; timeout 200 python bench.py --workload config2 --no-cpu-baseline --no-predict 2>&1 | tail -1 | python -c "$SUM"
ls -la gpurun_out | tail -3
