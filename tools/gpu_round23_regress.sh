#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r01r}
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.txt
echo "=== smoke"; timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1 | cut -c1-200
echo "=== bench default"; timeout 600 python bench.py 2>gpurun_out/${TAG}_bench.err | tail -1 | tee gpurun_out/${TAG}_bench.json | cut -c1-1500
echo "=== bench reference"; timeout 300 python bench.py --impl reference --steps 1 --warmup 1 2>&1 | tail -1 | cut -c1-300
