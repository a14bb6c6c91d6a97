import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from wisecondorx_b200 import cbs, synth
rng = np.random.default_rng(0)
per = synth.config_bins(3)
series = []
for c, n in enumerate(per):
    y = rng.normal(0, 0.1, int(n)); w = rng.uniform(0.5, 2, int(n))
    if c == 4: y[2000:2400] += 0.5
    series.append((y, w))
for rep in range(2):
    t0 = time.perf_counter()
    ends = cbs.segment_series(series, nperm=10000, seed=1)
    print("wall", time.perf_counter() - t0, cbs.cbs_stats(), [len(e) for e in ends][:6])
