#!/bin/bash
mkdir -p gpurun_out
nproc; free -g | head -2
echo "=== cli tests"; timeout 600 python -m pytest tests/test_cli_gpu.py tests/test_prep_gpu.py -q -x --tb=short 2>&1 | tail -5
echo "=== newref CLI wall-clock, config 2 (100 x 100 kb)"; timeout 600 python tools/newref_cli_wallclock.py --samples 100 --binsize 100000 --predict 2>&1 | tail -1 | tee gpurun_out/r01m_cli_config2.json
echo "=== newref CLI wall-clock, config 3 (500 x 15 kb)"; timeout 1500 python tools/newref_cli_wallclock.py --samples 500 --binsize 15000 --predict 2>&1 | tail -3 | tee gpurun_out/r01m_cli_config3.json
