#!/bin/bash
# round-15: CBS shared-memory shuffle; whole GPU suite
mkdir -p gpurun_out
TAG=${1:-r01k}
SUM='import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ["value","ms_per_step","stages_ms","exact_fallback_rows"]}, d["e2e"]["ms_per_step"], d["roofline"]["frac"]); print(json.dumps(d.get("predict")))'
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -8
echo "=== bench config3"; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench.json | python -c "$SUM"
echo "=== predict launch list"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}_predict.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/launches_${TAG}_predict.log 2>&1
ls -la gpurun_out | tail -3
