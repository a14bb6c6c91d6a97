#!/bin/bash
mkdir -p gpurun_out
SUM='import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ["value","ms_per_step","stages_ms","exact_fallback_rows"]}, d["e2e"]["ms_per_step"], d["roofline"]["frac"])'
echo "=== newref gpu tests"; timeout 280 python -m pytest tests/test_newref_gpu.py -q -x --tb=short 2>&1 | tail -5
echo "=== bench config3 kernel=5"; WCX_DEBUG_FAIL=1 timeout 120 python bench.py --kernel 5 --steps 2 --warmup 1 --no-cpu-baseline --no-predict 2>gpurun_out/fail.txt | tail -1 | python -c "$SUM"; sort gpurun_out/fail.txt | uniq -c | head
echo "=== list stats"; timeout 120 python - <<'PY'
import numpy as np
from wisecondorx_b200 import _lib, newref_tools
import bench
x, per, cum = bench.make_workload("config3")
eng = newref_tools.NewrefEngine(0)
eng.load(x, per, cum)
n = x.shape[0]
idx, dist = eng.topk(0, n, 300)
st = eng.stats(); print(st, eng.stage_ms())
nl = st["column_splits"] * 2
cnt = np.zeros(n * nl, dtype=np.int32)
_lib.check(_lib.load().wcx_debug_list_counts(eng.ctx.handle, cnt.ctypes.data, cnt.size))
cnt = cnt.reshape(-1, nl)
print("list length per list: mean %.1f max %d min %d ; per row total mean %.1f" % (cnt.mean(), cnt.max(), cnt.min(), cnt.sum(1).mean()))
PY
