#!/bin/bash
mkdir -p gpurun_out
SUM='import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ["value","ms_per_step","stages_ms","exact_fallback_rows"]}, d["e2e"]["ms_per_step"])'
echo "=== all gpu tests"; timeout 280 python -m pytest tests -q -m gpu -x --tb=short 2>&1 | tail -15
echo "=== bench config3"; timeout 120 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-predict 2>&1 | tail -1 | python -c "$SUM"
echo "=== bench config2"; timeout 120 python bench.py --workload config2 --steps 3 --warmup 2 --no-cpu-baseline --no-predict 2>&1 | tail -1 | python -c "$SUM"
echo "=== ncu counters"
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"rerank_kernel|null_ratios_kernel" -c 3 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-predict 2>&1 | grep -E "rerank_kernel|null_ratios_kernel|gpu__time|inst_executed|wavefronts|issue_active|lts__" | head -40
