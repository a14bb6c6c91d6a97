#!/bin/bash
mkdir -p gpurun_out
echo "=== prep + cli tests"; timeout 600 python -m pytest tests/test_prep_gpu.py tests/test_cli_gpu.py -q -x --tb=short 2>&1 | tail -5
echo "=== CLI config3"; timeout 900 python tools/newref_cli_wallclock.py --samples 500 --binsize 15000 --predict 2>&1 | tail -1 | tee gpurun_out/r01p_cli_config3.json
