#!/bin/bash
# First GPU session: SIMT/exact parity, then the tcgen05 kernel in isolation, then everything, then bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep "Model name" >> gpurun_out/gpu.txt
echo "=== simt/exact goldens"; timeout 600 python -m pytest tests/test_newref_gpu.py -x -q -k "prep or (golden and (simt or exact)) or random_draw" 2>&1 | tail -15
echo "=== tc tile"; timeout 300 python -m pytest tests/test_newref_gpu.py -x -q -k "tensor_core_tile" 2>&1 | tail -25
echo "=== tc goldens"; timeout 300 python -m pytest tests/test_newref_gpu.py -x -q -k "golden and tc" 2>&1 | tail -15
echo "=== all gpu tests"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25
echo "=== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5
echo "=== bench config2"; timeout 600 python bench.py --workload config2 --steps 3 --warmup 1 --cpu-seconds 5 2>&1 | tail -3
echo "=== bench config3"; timeout 1200 python bench.py --steps 3 --warmup 1 --cpu-seconds 10 2>&1 | tail -3
