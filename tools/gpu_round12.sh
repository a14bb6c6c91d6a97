#!/bin/bash
# round-12: CBS permutation statistic split into prep / arcs / count, t-test preparation batched per round
mkdir -p gpurun_out
TAG=${1:-r01h}
SUM='import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ["value","ms_per_step","stages_ms","exact_fallback_rows"]}, d["e2e"]["ms_per_step"], d["roofline"]["frac"]); print(json.dumps(d.get("predict")))'
echo "=== gpu tests (cbs, predict, cli)"; timeout 900 python -m pytest tests/test_cbs_gpu.py tests/test_predict_gpu.py tests/test_cli_gpu.py -q -x --tb=short 2>&1 | tail -15
echo "=== bench config3"; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench.json | python -c "$SUM"
echo "=== predict launch list"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}_predict.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/launches_${TAG}_predict.log 2>&1
ls -la gpurun_out | tail -3
