#!/bin/bash
# One GPU-box session: tests, smoke, bench, launch list, optional full captures.  Usage: tools/gpu_run.sh TAG [steps...]
# steps: tests smoke bench launches ncu_newref ncu_predict bench2 (default: tests smoke bench)
mkdir -p gpurun_out
TAG=${1:-r02a}; shift
STEPS=${@:-tests smoke bench}
for S in $STEPS; do
  case $S in
    tests) echo "=== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.txt ;;
    smoke) echo "=== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1 | cut -c1-300 ;;
    bench) echo "=== bench"; timeout 900 python bench.py 2>gpurun_out/${TAG}_bench.err | tail -1 | tee gpurun_out/${TAG}_bench.json | cut -c1-1500 ;;
    sweeptest) echo "=== sweep variants (short, guarded by a 150 s limit)"
      timeout 150 python -m pytest tests/test_newref_gpu.py -m gpu -x -q -k "streaming or config2 or edge" </dev/null 2>&1 | tail -5
      [ ${PIPESTATUS[0]} -ne 0 ] && { echo "sweeptest failed: stopping"; exit 1; } ;;
    benchq|benchq_own) echo "=== quick bench ($S): 10 steps, no predict / CLI / CPU baseline"
      ENV=""; [ $S == benchq_own ] && ENV="WCX_SWEEP_MULTIPLY_OWN=1"
      env $ENV timeout 300 python bench.py --steps 10 --warmup 3 --no-predict --no-cpu-baseline </dev/null 2>gpurun_out/${TAG}_${S}.err | tail -1 > gpurun_out/${TAG}_${S}.json
      python - <<PYEOF
import json
try:
    d = json.load(open("gpurun_out/${TAG}_${S}.json"))
    print("$S ms/step %.2f e2e %.2f frac %.3f sweep kernel %.2f ms" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms"]), {k: round(v, 2) for k, v in d["stages_ms"].items()}, d.get("parity"), d.get("counters"))
except Exception as e:
    print("$S failed:", e); print(open("gpurun_out/${TAG}_${S}.err").read()[-1500:])
PYEOF
      ;;
    cli|cli_prof) echo "=== command line at config 3 ($S)"
      EXTRA=""; [ $S == cli_prof ] && EXTRA="--cprofile"
      timeout 400 python tools/newref_cli_wallclock.py --predict $EXTRA </dev/null 2>gpurun_out/${TAG}_${S}.err | tail -1 | tee gpurun_out/${TAG}_${S}.json | cut -c1-900
      [ $S == cli_prof ] && grep -A50 "cumulative" gpurun_out/${TAG}_${S}.err | cut -c1-150 | head -60 ;;
    benchref) echo "=== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/${TAG}_bench.err | tail -1 | tee gpurun_out/${TAG}_bench_reference.json | cut -c1-600 ;;
    launches) echo "=== ncu launch list"
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
      tail -1 gpurun_out/${TAG}_launches.log | cut -c1-300 ;;
    ncu_newref)
      for K in dist_topk_tc rerank null_fast; do
        timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -f -o gpurun_out/${TAG}_prof_$K python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-predict > gpurun_out/${TAG}_prof_$K.log 2>&1
        python tools/ncu_summary.py gpurun_out/${TAG}_prof_$K.ncu-rep > gpurun_out/${TAG}_ncu_$K.txt 2>&1
      done ;;
    ncu_predict)
      timeout 900 ncu --set full --clock-control none -k 'regex:normalize_pass|nanmedian|radix_|weights_kernel|cutoff_partial|segment_z|cbs_prepare|cbs_tailp|cbs_maxarc|coverage_gather|project_apply' -c 48 -f -o gpurun_out/${TAG}_prof_predict python tools/predict_profile.py > gpurun_out/${TAG}_prof_predict.log 2>&1
      python tools/ncu_summary.py gpurun_out/${TAG}_prof_predict.ncu-rep > gpurun_out/${TAG}_ncu_predict.txt 2>&1
      [ -f gpurun_out/${TAG}_prof_predict.ncu-rep ] && [ $(stat -c %s gpurun_out/${TAG}_prof_predict.ncu-rep) -gt 30000000 ] && rm -f gpurun_out/${TAG}_prof_predict.ncu-rep  # gpurun_out is capped at 64 MiB
      tail -3 gpurun_out/${TAG}_prof_predict.log | cut -c1-300 ;;
    bench2|bench4|bench8) N=${S#bench}; echo "=== bench on $N GPUs"
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/${TAG}_bench_${N}gpu.err | tail -1 | tee gpurun_out/${TAG}_bench_${N}gpu.json | cut -c1-300
      python - <<PYEOF
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_${N}gpu.json"))
    print("N=$N ms/step %.2f e2e %.2f frac %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"]), {k: round(v, 2) for k, v in d["stages_ms"].items()}, d["parity"], d.get("predict"))
except Exception as e:
    print("bench $N failed:", e); print(open("gpurun_out/${TAG}_bench_${N}gpu.err").read()[-1500:])
PYEOF
      ;;
    ncu_null) echo "=== ncu: null-ratio kernels (stand-alone call)"
      timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:null_fast|null_ratios_kernel|nq_|gather_cols' -c 8 -f -o gpurun_out/${TAG}_prof_null python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-predict --unfused > gpurun_out/${TAG}_prof_null.log 2>&1
      python tools/ncu_summary.py gpurun_out/${TAG}_prof_null.ncu-rep > gpurun_out/${TAG}_ncu_null.txt 2>&1
      python tools/ncu_table.py gpurun_out/${TAG}_ncu_null.txt
      ncu -i gpurun_out/${TAG}_prof_null.ncu-rep --page details --csv 2>/dev/null | grep -i -E "null_fast" | grep -i -E "Issue Slots Busy|Executed Ipc|No Eligible|Stall|Theoretical Occupancy|Achieved Occupancy|L1/TEX Hit|Local" | cut -d, -f5,13- | head -40
      ;;
    nullcmp) echo "=== null ratios: side stream vs serial vs stand-alone"
      for MODE in side serial unfused; do
        EXTRA=""; ENV=""
        [ $MODE == serial ] && ENV="WCX_SERIAL_NULLS=1"
        [ $MODE == unfused ] && EXTRA="--unfused"
        env $ENV timeout 300 python bench.py --steps 10 --warmup 3 --no-predict --no-cpu-baseline $EXTRA 2>/dev/null | tail -1 > gpurun_out/${TAG}_bench_$MODE.json
        python - <<PYEOF
import json
d = json.load(open("gpurun_out/${TAG}_bench_$MODE.json"))
print("$MODE", "ms/step %.2f" % d["ms_per_step"], "e2e %.2f" % d["e2e"]["ms_per_step"], {k: round(v, 2) for k, v in d["stages_ms"].items()}, d["parity"]["ok"])
PYEOF
      done ;;
    prof_batch) echo "=== predict batch 96: host profile + CBS launch list"
      timeout 600 python tools/predict_profile.py --batch 96 --cprofile > gpurun_out/${TAG}_batch96_cprofile.txt 2>&1
      head -c 3000 gpurun_out/${TAG}_batch96_cprofile.txt | tail -c 1500
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:cbs_|segment_z' --csv --log-file gpurun_out/${TAG}_batch96_cbs_launches.csv python tools/predict_profile.py --batch 96 > gpurun_out/${TAG}_batch96_cbs_launches.log 2>&1
      python - <<PYEOF
import csv, collections
rows = list(csv.reader(open("gpurun_out/${TAG}_batch96_cbs_launches.csv")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
ki, vi = rows[hdr].index("Kernel Name"), rows[hdr].index("Metric Value")
acc = collections.Counter(); cnt = collections.Counter()
for r in rows[hdr + 2:]:
    if len(r) > vi:
        n = r[ki].split("(")[0][:60]; acc[n] += float(r[vi].replace(",", "")); cnt[n] += 1
for n, v in acc.most_common(): print("%-60s %6d launches %10.3f ms" % (n, cnt[n], v / 1e6))
PYEOF
      ;;
    *) echo "=== custom: $S"; timeout 300 bash -c "$S" </dev/null 2>&1 | tail -30 ;;
  esac
done
du -sh gpurun_out; ls -la gpurun_out | tail -20
