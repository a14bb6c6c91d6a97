#!/bin/bash
# round-13: null ratios fused into the re-rank kernel, blocked D2H overlap
mkdir -p gpurun_out
TAG=${1:-r01i}
SUM='import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ["value","ms_per_step","stages_ms","exact_fallback_rows"]}, d["e2e"]["ms_per_step"], d["roofline"]["frac"])'
echo "=== gpu tests (newref)"; timeout 900 python -m pytest tests/test_newref_gpu.py -q -x --tb=short 2>&1 | tail -15
echo "=== smoke"; timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "=== bench config3 fused"; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-predict 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench.json | python -c "$SUM"
echo "=== bench config3 unfused"; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-predict --unfused 2>&1 | tail -1 | python -c "$SUM"
echo "=== bench config2 fused"; timeout 200 python bench.py --workload config2 --no-cpu-baseline --no-predict 2>&1 | tail -1 | python -c "$SUM"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rerank_kernel -c 1 -o gpurun_out/prof_${TAG}_rerank_fused python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-predict > gpurun_out/prof_${TAG}_rerank.log 2>&1
ls -la gpurun_out | tail -3
