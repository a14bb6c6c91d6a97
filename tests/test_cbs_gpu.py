"""GPU tests of the CUDA circular binary segmentation through the C-ABI: breakpoints identical to
the CPU restatement (oracle/cbs_oracle.py; same Philox streams and arithmetic order), planted
breakpoints recovered, nothing on noise, and the reference's example (chr21 gain).  DNAcopy itself
is PARITY UNPINNED (see oracle/cbs_oracle.py)."""
import os
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import cbs_oracle as C  # noqa: E402
from wisecondorx_b200 import cbs  # noqa: E402


def _series(rng, n, steps, sd=0.05):
    y = rng.normal(0, sd, n)
    for (a, b, h) in steps:
        y[a:b] += h
    return y, rng.uniform(0.4, 2.5, n)


def test_breakpoints_equal_oracle_batch():
    rng = np.random.default_rng(11)
    specs = [(600, [(200, 260, 0.25), (400, 600, -0.1)]), (150, [(50, 90, 0.2)]), (1500, [(700, 705, 0.6), (1000, 1500, 0.08)]),
             (90, []), (40, [(10, 30, 0.5)]), (3, []), (2500, [(0, 40, -0.15), (2400, 2500, 0.12)]), (300, [(100, 108, 0.12)])]
    series = [_series(rng, n, st) for n, st in specs]
    got = cbs.segment_series(series, alpha=1e-4, nperm=300, seed=5)
    for c, ((y, w), g) in enumerate(zip(series, got)):
        want = C.segment_chromosome(y, w, alpha=1e-4, nperm=300, seed=5, chrom=c)
        assert g.tolist() == want, (c, g.tolist(), want)
    st = cbs.cbs_stats()
    assert st["segments_tested"] > len(series)


def test_permutation_paths_equal_oracle():
    """Borderline effects so that the hybrid / full permutation tests and the edge t-tests decide."""
    rng = np.random.default_rng(12)
    series = []
    for n, a, b, h in [(260, 100, 140, 0.045), (260, 100, 140, 0.06), (150, 60, 75, 0.07), (150, 60, 75, 0.1), (700, 300, 330, 0.05),
                       (700, 300, 330, 0.065), (120, 5, 20, 0.09), (500, 200, 212, 0.09)]:
        series.append(_series(rng, n, [(a, b, h)]))
    got = cbs.segment_series(series, alpha=1e-3, nperm=400, seed=9)
    st = cbs.cbs_stats()
    assert st["perm_tests"] > 0
    for c, ((y, w), g) in enumerate(zip(series, got)):
        want = C.segment_chromosome(y, w, alpha=1e-3, nperm=400, seed=9, chrom=c)
        assert g.tolist() == want, (c, g.tolist(), want)


def test_sequential_boundary_rule_equals_oracle():
    """DNAcopy's early-stopping boundary (wcx_cbs_set_boundary) with several tolerated exceedances (alpha * nperm = 8:
    rows 1 .. 9 of the table) and without it: the same decisions as the sequential loop of the oracle, and the two
    rules are allowed to differ from each other only by construction."""
    rng = np.random.default_rng(14)
    series = []
    for n, a, b, h in [(180, 60, 100, 0.03), (180, 60, 100, 0.04), (150, 60, 75, 0.05), (150, 60, 75, 0.06), (190, 20, 60, 0.035),
                       (120, 5, 20, 0.06), (199, 90, 110, 0.05), (500, 200, 230, 0.04), (900, 300, 340, 0.03)]:
        series.append(_series(rng, n, [(a, b, h)]))
    for sequential in (True, False):
        got = cbs.segment_series(series, alpha=0.02, nperm=400, seed=3, sequential=sequential)
        assert cbs.cbs_stats()["perm_tests"] > 0
        for c, ((y, w), g) in enumerate(zip(series, got)):
            want = C.segment_chromosome(y, w, alpha=0.02, nperm=400, seed=3, chrom=c, sequential=sequential)
            assert g.tolist() == want, (sequential, c, g.tolist(), want)


def test_noise_and_planted():
    rng = np.random.default_rng(13)
    noise = [_series(rng, n, []) for n in (60, 190, 1000, 4000)]
    for (y, w), g in zip(noise, cbs.segment_series(noise, nperm=1000, seed=1)):
        assert g.tolist() == [len(y)]
    y, w = _series(rng, 16000, [(3000, 3400, 0.05), (9000, 16000, -0.02)], sd=0.03)
    g = cbs.segment_series([(y, w)], nperm=1000, seed=1)[0].tolist()
    assert len(g) == 4 and g[3] == 16000
    assert all(abs(a - b) <= 3 for a, b in zip(g, [3000, 3400, 9000, 16000])), g
    assert g == C.segment_chromosome(y, w, nperm=1000, seed=1, chrom=0)


def test_exec_cbs_flow_equals_oracle_on_example_bed(golden_dir):
    g = np.load(os.path.join(golden_dir, "example_bed.npz"))
    chrs, ratio = g["chr"].astype(int), g["ratio"].astype(np.float64)
    rng = np.random.default_rng(3)
    rr = [ratio[chrs == c].tolist() for c in range(1, 24)]
    ww = [rng.uniform(0.5, 2.0, int((chrs == c).sum())).tolist() for c in range(1, 24)]
    got = cbs.cbs_segments(rr, ww, "F", 1e-4, 100000, seed=7, nperm=200)
    want = C.exec_cbs_segments(rr, ww, "F", 1e-4, 100000, seed=7, nperm=200)
    assert [s[:3] for s in got] == [s[:3] for s in want]
    np.testing.assert_allclose([s[3] for s in got], [s[3] for s in want], rtol=1e-14)
    c21 = max((s for s in got if s[0] == 20), key=lambda s: s[2] - s[1])
    assert abs(c21[1] - 131) <= 3 and c21[2] == 467  # random weights may move the left edge by a bin or two


def test_exec_cbs_dropin_signature():
    """exec_cbs(rem_input, results) -> [[chr, s, e, z, r], ...] like predict_tools.exec_cbs."""
    rng = np.random.default_rng(4)
    bpc = [120, 90] + [30] * 21
    rr, ww, nrs = [], [], []
    for c, nb in enumerate(bpc):
        r = rng.normal(0, 0.03, nb)
        if c == 1:
            r[20:60] += 0.3
        r[rng.random(nb) < 0.05] = 0
        rr.append(r.tolist()); ww.append(rng.uniform(0.5, 2, nb).tolist())
        nrs.append([rng.normal(0, 0.03, 25).tolist() for _ in range(nb)])
    rem = {"args": types.SimpleNamespace(alpha=1e-4, seed=None), "ref_gender": "F", "binsize": 100000}
    res = {"results_r": rr, "results_w": ww, "results_nr": nrs}
    out = cbs.exec_cbs(rem, res, nperm=500)
    assert all(len(row) == 5 for row in out)
    gain = [row for row in out if row[0] == 1 and row[4] > 0.2]
    assert len(gain) == 1 and abs(gain[0][1] - 20) <= 1 and abs(gain[0][2] - 60) <= 1 and gain[0][3] > 5


def test_boundary_arguments_and_reset():
    """wcx_cbs_set_boundary: the table stays in the context until replaced; n = 0 switches the rule off; values below 1
    and a NULL table with n > 0 are rejected."""
    import ctypes

    from wisecondorx_b200 import _lib
    ctx = _lib.default_context(0)
    L = _lib.load()
    bad = np.array([9500, 0, 9864], dtype=np.int32)
    assert L.wcx_cbs_set_boundary(ctx.handle, ctypes.c_void_p(bad.ctypes.data), 3) != 0
    assert L.wcx_cbs_set_boundary(ctx.handle, None, 3) != 0
    rng = np.random.default_rng(15)
    series = [_series(rng, 260, [(100, 140, 0.05)]), _series(rng, 150, [(60, 75, 0.08)])]
    a = cbs.segment_series(series, alpha=1e-3, nperm=400, seed=2, sequential=False)
    # a table whose every entry is 1: any test that reaches the permutations is declared significant after the first one
    ones = np.ones(3, dtype=np.int32)
    _lib.check(L.wcx_cbs_set_boundary(ctx.handle, ctypes.c_void_p(ones.ctypes.data), 3))
    # segment_series sets its own table (or none) on every call: the leftover table above must not leak into it
    b = cbs.segment_series(series, alpha=1e-3, nperm=400, seed=2, sequential=False)
    assert [x.tolist() for x in a] == [x.tolist() for x in b]


def test_pinned_pool_reserve_and_prewarm():
    """_lib.pinned: reserve() leaves a page-locked buffer that a slightly smaller request reuses (tool_newref reserves
    upper bounds on a background thread), a much smaller request does not; prewarm_async is safe to call twice."""
    from wisecondorx_b200 import _lib
    t = _lib.prewarm_async(0, [6 << 20])
    t.join()
    _lib.prewarm_async(0).join()
    pool = _lib.pinned
    have = sum(len(v) for k, v in pool._free.items() if k == 6 << 20)
    assert have >= 1
    a = pool.empty((5 << 20,), np.uint8)  # 5 MiB fits the reserved 6 MiB buffer (within the 1.25 slack)
    assert sum(len(v) for k, v in pool._free.items() if k == 6 << 20) == have - 1
    a[:] = 7
    small = pool.empty((1 << 20,), np.uint8)  # 1 MiB must not take a 6 MiB buffer
    del a
    assert sum(len(v) for k, v in pool._free.items() if k == 6 << 20) == have
    del small
