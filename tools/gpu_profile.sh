#!/bin/bash
# ncu evidence for the current build: launch list + one full capture per main kernel
mkdir -p gpurun_out
TAG=${1:-r01b}
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_${TAG}.log 2>&1
for K in dist_topk_tc rerank_kernel null_ratios_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -o gpurun_out/prof_${TAG}_$K python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/prof_${TAG}_$K.log 2>&1
done
ls -la gpurun_out
