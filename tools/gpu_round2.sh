#!/bin/bash
mkdir -p gpurun_out
echo "=== tc tile + goldens"; timeout 600 python -m pytest tests/test_newref_gpu.py -x -q -k "tensor_core_tile or golden" 2>&1 | tail -15
echo "=== all gpu tests"; timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -60
echo "=== list stats"; timeout 600 python - <<'PY'
import numpy as np, ctypes
from wisecondorx_b200 import _lib, newref_tools, synth
import bench
x, per, cum = bench.make_workload("config3")
eng = newref_tools.NewrefEngine(0)
eng.load(x, per, cum)
n = x.shape[0]
rb, re = 0, n
idx, dist = eng.topk(rb, re, 300)
print(eng.stats(), eng.stage_ms())
st = eng.stats()
nl = st["column_splits"] * 2
cnt = np.zeros((re - rb) * nl, dtype=np.int32)
_lib.check(_lib.load().wcx_debug_list_counts(eng.ctx.handle, cnt.ctypes.data, cnt.size))
cnt = cnt.reshape(-1, nl)
print("list length per list: mean %.1f max %d min %d ; per row total mean %.1f" % (cnt.mean(), cnt.max(), cnt.min(), cnt.sum(1).mean()))
PY
echo "=== bench config2"; timeout 600 python bench.py --workload config2 --steps 3 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ['value','ms_per_step','stages_ms','exact_fallback_rows']}, d['e2e'], d['roofline'])"
echo "=== bench config3"; timeout 1200 python bench.py --steps 3 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ['value','ms_per_step','stages_ms','exact_fallback_rows']}, d['e2e'], d['roofline'])"
