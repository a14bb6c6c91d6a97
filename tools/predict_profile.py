#!/usr/bin/env python
"""One predict pass (normalize for 1 sample, CBS, segment z-scores) against a config-3-size reference built on the
GPU -- the workload the ncu captures of the predict / CBS kernels are taken on (tools/gpu_run.sh ncu_predict)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from wisecondorx_b200 import _lib, newref_tools  # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "config3"
    x, per, cum = bench.make_workload(workload)
    n, s = x.shape
    dev = torch.device("cuda", 0)
    eng = newref_tools.NewrefEngine(0, _lib.Context(0))
    k, m = bench.REFSIZE, min(s, bench.NULL_M)
    idx = torch.empty((n, k), dtype=torch.int32, device=dev)
    dist = torch.empty((n, k), dtype=torch.float64, device=dev)
    nr = torch.empty((n, m), dtype=torch.float64, device=dev)
    eng.load(x, per, cum)
    eng.reference(0, n, k, np.arange(m, dtype=np.int32), device_out=(idx.data_ptr(), dist.data_ptr(), nr.data_ptr()))
    torch.cuda.synchronize()
    out = bench.predict_extras(eng, x, per, cum, idx, dist, nr, 0, batches=(1,))
    print(out)


if __name__ == "__main__":
    main()
