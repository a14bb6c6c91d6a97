"""Deterministic stand-in for the R bridge (exec_R -> CBS.R -> DNAcopy) used where the assembly around it is what
is under test (R is absent, `tool_test` runs with exec_R stubbed).  The same function feeds the live
reference in tests/golden/make_golden.py and the repo's pipeline in tests/test_assembly_gpu.py."""
import numpy as np


def stub_segments(results_r, results_w, nchr):
    """Two segments per chromosome ([0, n/2) and [n/2, n)), ratio = weighted mean of the bins with data;
    halves without data are dropped (as CBS.R drops all-NA chromosomes).  Returns [[chr0, s, e, r], ...]."""
    out = []
    for c in range(nchr):
        r = np.asarray(results_r[c], dtype=float)
        w = np.asarray(results_w[c], dtype=float)
        n = len(r)
        for s, e in ((0, n // 2), (n // 2, n)):
            m = r[s:e] != 0
            if e - s < 2 or not m.any():
                continue
            out.append([c, s, e, float(np.sum(r[s:e][m] * w[s:e][m]) / np.sum(w[s:e][m]))])
    return out
