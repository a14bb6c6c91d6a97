#!/bin/bash
# final round-1 evidence: full GPU suite, smoke, both bench arms, CLI wall-clock, ncu launch list + full captures
mkdir -p gpurun_out
TAG=${1:-r01n}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_smi.txt
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest.txt
echo "=== smoke"; timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "=== bench reference"; timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_ref.json | cut -c1-400
echo "=== bench default"; timeout 600 python bench.py 2>gpurun_out/${TAG}_bench.err | tail -1 | tee gpurun_out/${TAG}_bench.json
echo "=== bench config2"; timeout 200 python bench.py --workload config2 --no-cpu-baseline --no-predict 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_c2.json | cut -c1-600
echo "=== CLI config3"; timeout 900 python tools/newref_cli_wallclock.py --samples 500 --binsize 15000 --predict 2>&1 | tail -1 | tee gpurun_out/${TAG}_cli_config3.json
echo "=== ncu launch list"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-predict > gpurun_out/launches_${TAG}.log 2>&1
for K in dist_topk_tc rerank_kernel null_ratios_kernel cbs_maxarc; do
  EXTRA="--no-predict"; [ "$K" = "cbs_maxarc" ] && EXTRA=""
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -o gpurun_out/prof_${TAG}_$K python bench.py --steps 1 --warmup 0 --no-cpu-baseline $EXTRA > gpurun_out/prof_${TAG}_$K.log 2>&1
done
ls -la gpurun_out | tail -12
