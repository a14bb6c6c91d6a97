#!/bin/bash
# round-11: re-rank variants (pair / single candidate per lane group), batched CBS permutation tests
mkdir -p gpurun_out
TAG=${1:-r01g}
SUM='import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ["value","ms_per_step","stages_ms","exact_fallback_rows"]}, d["e2e"]["ms_per_step"], d["roofline"]["frac"]); print(json.dumps(d.get("predict")))'
echo "=== gpu tests (newref, cbs, predict)"; timeout 900 python -m pytest tests/test_newref_gpu.py tests/test_cbs_gpu.py tests/test_predict_gpu.py -q -x --tb=short 2>&1 | tail -15
echo "=== bench config3 pair"; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench.json | python -c "$SUM"
echo "=== bench config3 single"; WCX_RERANK_PAIR=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-predict 2>&1 | tail -1 | python -c "$SUM"
echo "=== bench config2 pair"; timeout 200 python bench.py --workload config2 --no-cpu-baseline --no-predict 2>&1 | tail -1 | python -c "$SUM"
echo "=== bench config2 single"; WCX_RERANK_PAIR=0 timeout 200 python bench.py --workload config2 --no-cpu-baseline --no-predict 2>&1 | tail -1 | python -c "$SUM"
WCX_RERANK_PAIR=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:rerank_kernel -c 1 -o gpurun_out/prof_${TAG}_rerank_single python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-predict > gpurun_out/prof_${TAG}_rerank.log 2>&1
echo "=== predict launch list"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}_predict.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/launches_${TAG}_predict.log 2>&1
ls -la gpurun_out | tail -4
