#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r01s}
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.txt
echo "=== smoke"; timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1 | cut -c1-120
echo "=== bench default"; timeout 600 python bench.py 2>gpurun_out/${TAG}_bench.err | tail -1 | tee gpurun_out/${TAG}_bench.json | cut -c1-400
echo "=== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-predict > gpurun_out/launches_${TAG}.log 2>&1
tail -2 gpurun_out/launches_${TAG}.log | cut -c1-300
