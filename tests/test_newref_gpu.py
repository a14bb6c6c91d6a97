"""GPU parity tests for the newref path (get_reference): the CUDA library, called through the
C-ABI (ctypes), against the oracle and the golden vectors of the live reference.

Bars (BASELINE.json north_star): indexes bit-exact; distances / null ratios within 1e-5 -- the
distance tests below assert the stronger bit-exact equality, which the exact re-rank guarantees.
"""
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import c_oracle, np_oracle  # noqa: E402
from wisecondorx_b200 import _lib, newref_tools, synth  # noqa: E402

# tc2h: the product path (tcgen05, CTA pairs, f16 operands); tch: the same kernel with one CTA per SM; simt: CUDA-core
# cross-check kernel; exact: brute-force float64 rows.  (The tf32 variants of round 1 are gone; their ids are aliases.)
KERNELS = {"simt": _lib.KERNEL_SIMT, "exact": _lib.KERNEL_EXACT, "tc2h": _lib.KERNEL_TC2H, "tch": _lib.KERNEL_TCH,
           "auto": _lib.KERNEL_AUTO}


@pytest.fixture(scope="module")
def gref(golden_dir):
    return np.load(os.path.join(golden_dir, "get_reference.npz"))


@pytest.fixture(scope="module")
def eng():
    return newref_tools.NewrefEngine(0)


def tf32_round(a):
    """round-to-nearest (ties away) to 10 mantissa bits, like cvt.rna.tf32.f32"""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x1000) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32)


def test_prep_centres_and_rounds(eng):
    x, per, cum = synth.make_corrected_matrix([300, 200, 100] + [10] * 19, 37, seed=3)
    eng.load(x, per, cum)
    import ctypes
    n, s = x.shape
    kp = ctypes.c_int32()
    L = _lib.load()
    _lib.check(L.wcx_debug_prep(eng.ctx.handle, None, None, ctypes.byref(kp)))
    assert kp.value == 64
    xc = np.empty((n, kp.value), dtype=np.float32)
    nrm = np.empty(n, dtype=np.float32)
    _lib.check(L.wcx_debug_prep(eng.ctx.handle, xc.ctypes.data, nrm.ctypes.data, None))
    want = tf32_round((x - x.mean(axis=0)).astype(np.float32))
    assert np.array_equal(xc[:, :s], want)
    assert not xc[:, s:].any()
    np.testing.assert_allclose(nrm, (want.astype(np.float64) ** 2).sum(1), rtol=1e-6)


def _prep_f16(eng, n):
    import ctypes
    L = _lib.load()
    kp = ctypes.c_int32()
    _lib.check(L.wcx_debug_prep_f16(eng.ctx.handle, None, None, ctypes.byref(kp), None))
    xh = np.empty((n, kp.value), dtype=np.float16)
    nrm = np.empty(n, dtype=np.float32)
    sc = np.empty(2, dtype=np.float64)
    _lib.check(L.wcx_debug_prep_f16(eng.ctx.handle, xh.ctypes.data, nrm.ctypes.data, None, sc.ctypes.data))
    return xh, nrm, sc


def test_prep_f16_scales_centres_and_rounds(eng):
    """f16 operand set: (X - mean) * 2^e rounded once to nearest-even, |values| < 2^14, norms of the rounded values."""
    x, per, cum = synth.make_corrected_matrix([300, 200, 100] + [10] * 19, 37, seed=3)
    eng.load(x, per, cum)
    n, s = x.shape
    xh, nrm, sc = _prep_f16(eng, n)
    assert xh.shape[1] == 64
    bound = np.abs(x).max() + np.abs(x.mean(axis=0)).max()
    e = np.log2(sc[0])
    assert e == int(e) and sc[1] == sc[0] ** 2
    assert 2.0 ** 13 <= bound * sc[0] < 2.0 ** 14
    want = ((x - x.mean(axis=0)) * sc[0]).astype(np.float16)  # NumPy rounds float64 -> float16 to nearest even
    # the device mean (atomic accumulation order) may differ from NumPy's in the last bit: allow rare 1-ulp flips
    got = xh[:, :s]
    assert (got != want).mean() < 1e-3
    assert np.abs(got.astype(np.float64) - want.astype(np.float64)).max() <= np.spacing(np.abs(want).max())
    assert not xh[:, s:].any()
    np.testing.assert_allclose(nrm, (got.astype(np.float64) ** 2).sum(1), rtol=1e-6)


def test_tensor_core_tile_f16_matches_fp64_matmul(eng):
    """kind::f16 accumulators vs a float64 product of the same f16 operands: products of two 11-bit significands are
    exact in fp32, so only the accumulation may round -- the bound rerank.cu assumes is (K + 64) * 2^-23 * |a| |b|."""
    x, per, cum = synth.make_corrected_matrix([500, 400, 300] + [20] * 19, 100, seed=4)
    eng.load(x, per, cum)
    L = _lib.load()
    n, s = x.shape
    xh, nrm, sc = _prep_f16(eng, n)
    kp = xh.shape[1]
    worst = 0.0
    for row0, col0 in [(0, 0), (128, 256), (900, 1024)]:
        acc = np.empty((128, 256), dtype=np.float32)
        _lib.check(L.wcx_debug_tc_tile_f16(eng.ctx.handle, row0, col0, acc.ctypes.data))
        a = np.zeros((128, kp)); b = np.zeros((256, kp))
        ra = xh[row0:row0 + 128].astype(np.float64); rb = xh[col0:col0 + 256].astype(np.float64)
        a[:len(ra)] = ra; b[:len(rb)] = rb
        want = a @ b.T
        allowed = (kp + 64) * 2.0 ** -23 * np.sqrt((a * a).sum(1))[:, None] * np.sqrt((b * b).sum(1))[None, :]
        ratio = np.abs(acc - want) / np.maximum(allowed, 1e-300)
        worst = max(worst, float(ratio.max()))
    assert worst < 0.25, worst  # measured accumulation error stays far inside the assumed bound


@pytest.mark.parametrize("kernel", ["auto", "tc2h", "tch", "simt", "exact"])
@pytest.mark.parametrize("case,part,parts,k", [("A_p11", 1, 1, 30), ("A_p23", 2, 3, 30), ("G", 1, 1, 30),
                                                ("T", 1, 1, 12), ("S", 1, 1, 20)])
def test_golden_get_reference(gref, kernel, case, part, parts, k):
    base = case.split("_")[0]
    x, per, cum = gref[base + "_x"], gref[base + "_per"], gref[base + "_cum"]
    idx, dist, nr = newref_tools.get_reference(x, per, cum, k, part, parts, kernel=KERNELS[kernel],
                                               sample_ids=gref[case + "_ids"].tolist())
    assert idx.dtype == np.int32 and dist.dtype == np.float64
    assert np.array_equal(idx, gref[case + "_idx"])
    assert np.array_equal(dist, gref[case + "_dist"])
    np.testing.assert_allclose(nr, gref[case + "_nr"], rtol=1e-12, atol=1e-14, equal_nan=True)


def test_random_draw_follows_python_random(gref):
    """Without explicit sample_ids the wrapper draws like the reference (newref_tools.py:214-217)."""
    x, per, cum = gref["A_x"], gref["A_per"], gref["A_cum"]
    random.seed(7)
    idx, dist, nr = newref_tools.get_reference(x, per, cum, 30, 1, 1)
    np.testing.assert_allclose(nr, gref["A_p11_nr"], rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("kernel", ["tc2h", "tch", "simt"])
def test_config1_full_vs_c_oracle(eng, kernel):
    """BASELINE config 1: 1 Mb bins (2887 autosomal), 20 samples, refsize 300 -- full parity."""
    per = synth.config_bins(1)
    x, per, cum = synth.make_corrected_matrix(per, 20, seed=11)
    n = x.shape[0]
    eng.load(x, per, cum)
    idx, dist = eng.topk(0, n, 300, KERNELS[kernel])
    oi, od = c_oracle.topk(x, per, cum, 300, 0, n)
    assert np.array_equal(idx, oi)
    assert np.array_equal(dist, od)
    ids = list(range(20))
    nr = eng.null_ratios(0, n, 300, ids)
    onr = c_oracle.null_ratios(x, oi, 0, n, ids)
    np.testing.assert_allclose(nr, onr, rtol=1e-12, atol=1e-14)
    st = eng.stats()
    assert st["exact_fallback_rows"] <= n // 50, st


@pytest.mark.parametrize("kernel", ["tc2h", "tch", "simt"])
def test_config2_parts_vs_c_oracle(eng, kernel):
    """BASELINE config 2: 100 kb bins (28760), 100 samples, refsize 300; parts compared in full."""
    per = synth.config_bins(2)
    x, per, cum = synth.make_corrected_matrix(per, 100, seed=12)
    n = x.shape[0]
    eng.load(x, per, cum)
    parts = 100
    for part in (1, 37, 100):
        s, e = newref_tools._get_part(part - 1, parts, n)
        idx, dist = eng.topk(s, e, 300, KERNELS[kernel])
        oi, od = c_oracle.topk(x, per, cum, 300, s, e)
        assert np.array_equal(idx, oi), part
        assert np.array_equal(dist, od), part
        ids = list(range(0, 100, 7))
        nr = eng.null_ratios(s, e, 300, ids)
        onr = c_oracle.null_ratios(x, oi, s, e, ids)
        np.testing.assert_allclose(nr, onr, rtol=1e-12, atol=1e-14)
        assert eng.stats()["exact_fallback_rows"] <= (e - s) // 20


def test_tc_equals_simt_whole_config2(eng):
    per = synth.config_bins(2)
    x, per, cum = synth.make_corrected_matrix(per, 100, seed=13)
    n = x.shape[0]
    eng.load(x, per, cum)
    i1, d1 = eng.topk(0, n, 300, KERNELS["tc2h"])
    st1 = eng.stats()
    i2, d2 = eng.topk(0, n, 300, KERNELS["simt"])
    assert np.array_equal(i1, i2) and np.array_equal(d1, d2)
    i3, d3 = eng.topk(0, n, 300, KERNELS["tch"])
    assert np.array_equal(i1, i3) and np.array_equal(d1, d3)
    assert eng.stats()["exact_fallback_rows"] <= n // 100
    assert st1["exact_fallback_rows"] <= n // 100, st1
    # sortedness + index range: size-independent properties
    assert (np.diff(d1, axis=1) >= 0).all()
    assert i1.min() >= 0 and i1.max() < n


def test_edge_cases(eng):
    # ties everywhere (quantised data), NaN row, fewer candidates than refsize, empty range
    per = np.array([40, 30, 20, 10] + [0] * 18)
    x, per, cum = synth.make_corrected_matrix(per, 9, seed=21)
    x = np.round(x * 16) / 16
    x[5, 3] = np.nan
    n = x.shape[0]
    oi, od = c_oracle.topk(x, per, cum, 64, 0, n)
    for kernel in ("tc2h", "tch", "simt", "exact"):
        eng.load(x, per, cum)
        idx, dist = eng.topk(0, n, 64, KERNELS[kernel])
        assert np.array_equal(idx, oi), kernel
        assert np.array_equal(dist, od), kernel
        idx0, dist0 = eng.topk(7, 7, 64, KERNELS[kernel])
        assert idx0.shape == (0, 64)
    assert (oi[5] == -1).all() and (od[5] == 1e10).all()


def test_bad_arguments_raise(eng):
    x, per, cum = synth.make_corrected_matrix([50, 40] + [5] * 20, 8, seed=2)
    eng.load(x, per, cum)
    with pytest.raises(_lib.WcxError):
        eng.topk(0, x.shape[0] + 1, 10)
    with pytest.raises(_lib.WcxError):
        eng.topk(0, 10, 0)
    with pytest.raises(_lib.WcxError):
        eng.null_ratios(0, 10, 10, [99])
    n = x.shape[0]
    for bad in (n, n + 7, -n - 1):  # caller-supplied positions follow Python index semantics: [-n, n)
        idx = np.zeros((10, 10), dtype=np.int32)
        idx[3, 4] = bad
        with pytest.raises(_lib.WcxError):
            eng.null_ratios(0, 10, 10, [0, 1], idx=idx)
    idx = np.full((10, 10), -n, dtype=np.int32)  # -n is the first bin
    assert np.isfinite(eng.null_ratios(0, 10, 10, [0, 1], idx=idx)).all()


@pytest.mark.parametrize("kernel", ["tc2h", "tch"])
def test_nasty_data_vs_c_oracle(eng, kernel):
    """Duplicated bins (exact distance ties), outlier bins with huge norms, constant bins, a few
    NaN / inf rows, heavy-tailed noise: indexes and distances must still be bit-exact."""
    rng = np.random.default_rng(33)
    per = (synth.config_bins(2) // 4).astype(np.int64)
    x, per, cum = synth.make_corrected_matrix(per, 64, seed=34)
    n = x.shape[0]
    x *= 1.0 + 0.02 * rng.standard_t(2.5, size=x.shape)          # heavy tails
    dup = rng.choice(n, 400, replace=False)
    x[dup[200:]] = x[dup[:200]]                                   # 200 exact duplicate pairs
    x[rng.choice(n, 30, replace=False)] *= 25.0                   # outlier bins
    x[rng.choice(n, 20, replace=False)] = 1.0                     # constant bins (mutual distance 0)
    x[rng.choice(n, 3, replace=False), 5] = np.nan
    x[rng.choice(n, 2, replace=False), 7] = np.inf
    eng.load(x, per, cum)
    rows = np.concatenate([dup[:40], rng.choice(n, 200, replace=False)])
    idx, dist = eng.topk(0, n, 300, KERNELS[kernel])
    st = eng.stats()
    for r in np.sort(rows)[::7]:
        oi, od = c_oracle.topk(x, per, cum, 300, int(r), int(r) + 1)
        assert np.array_equal(idx[r], oi[0]), (kernel, int(r))
        assert np.array_equal(dist[r], od[0]), (kernel, int(r))
    assert st["exact_fallback_rows"] <= n // 20, st


@pytest.mark.parametrize("samples", [136, 777, 1100])
def test_many_samples_both_gather_paths(eng, samples):
    """Sample counts whose NumPy summation tree has 2 / 8 / 16 leaves: up to 8 leaves the re-rank gathers from the
    leaf-major copy of X (rerank.cu build_leaf_layout), beyond that it falls back to the row-major LDG gather."""
    per = [140, 120, 100, 90, 80, 70] + [25] * 16
    x, per, cum = synth.make_corrected_matrix(per, samples, seed=40 + samples)
    n = x.shape[0]
    eng.load(x, per, cum)
    idx, dist = eng.topk(0, n, 300)
    oi, od = c_oracle.topk(x, per, cum, 300, 0, n)
    assert np.array_equal(idx, oi)
    assert np.array_equal(dist, od)
    assert eng.stats()["exact_fallback_rows"] <= n // 20


def test_fused_reference_equals_separate_calls(eng):
    """wcx_newref_reference (null ratios fused into the re-rank kernel, rows in blocks) against topk + null_ratios
    and the C oracle; host and device-resident outputs; part ranges crossing chromosome borders."""
    import torch
    per = synth.config_bins(2)
    x, per, cum = synth.make_corrected_matrix(per, 100, seed=51)
    n = x.shape[0]
    eng.load(x, per, cum)
    ids = list(range(3, 100, 2)) + [0, 98]                    # 51 columns: a ragged last chunk of 3
    for s, e in ((0, 9000), (int(cum[2]) - 700, int(cum[2]) + 8000), (n - 300, n)):
        idx, dist, nr = eng.reference(s, e, 300, ids)
        oi, od = c_oracle.topk(x, per, cum, 300, s, e)
        onr = c_oracle.null_ratios(x, oi, s, e, ids)
        assert np.array_equal(idx, oi) and np.array_equal(dist, od)
        np.testing.assert_allclose(nr, onr, rtol=1e-12, atol=1e-14)
        i2, d2 = eng.topk(s, e, 300)
        n2 = eng.null_ratios(s, e, 300, ids)
        assert np.array_equal(idx, i2) and np.array_equal(dist, d2) and np.array_equal(nr, n2, equal_nan=True)
    # device-resident outputs
    s, e = 100, 4100
    ti = torch.empty((e - s, 300), dtype=torch.int32, device="cuda:0")
    td = torch.empty((e - s, 300), dtype=torch.float64, device="cuda:0")
    tn = torch.empty((e - s, len(ids)), dtype=torch.float64, device="cuda:0")
    eng.reference(s, e, 300, ids, device_out=(ti.data_ptr(), td.data_ptr(), tn.data_ptr()))
    eng.ctx.sync()
    idx, dist, nr = eng.reference(s, e, 300, ids)
    assert np.array_equal(ti.cpu().numpy(), idx) and np.array_equal(td.cpu().numpy(), dist)
    assert np.array_equal(tn.cpu().numpy(), nr, equal_nan=True)


def test_fused_reference_edge_cases(eng, gref):
    """Fewer candidates than ref_size (-1 fillers wrap to the last bin), NaN rows, gonosomal placeholder rows,
    large ref_size (unfused fall-back) -- against the oracle / golden vectors of the live reference."""
    per = np.array([40, 30, 20, 10] + [0] * 18)
    x, per, cum = synth.make_corrected_matrix(per, 9, seed=21)
    x[5, 3] = np.nan
    n = x.shape[0]
    ids = [0, 3, 8, 5]
    eng.load(x, per, cum)
    for k in (64, 350):
        idx, dist, nr = eng.reference(0, n, k, ids)
        oi, od = c_oracle.topk(x, per, cum, k, 0, n)
        onr = c_oracle.null_ratios(x, oi, 0, n, ids)
        assert np.array_equal(idx, oi) and np.array_equal(dist, od), k
        np.testing.assert_allclose(nr, onr, rtol=1e-12, atol=1e-14, equal_nan=True)
    # gonosomal reference of the golden file: autosomal rows are placeholders (idx 0, dist 1.0)
    gx, gper, gcum = gref["G_x"], gref["G_per"], gref["G_cum"]
    eng.load(gx, gper, gcum)
    idx, dist, nr = eng.reference(0, gx.shape[0], 30, gref["G_ids"].tolist())
    assert np.array_equal(idx, gref["G_idx"]) and np.array_equal(dist, gref["G_dist"])
    np.testing.assert_allclose(nr, gref["G_nr"], rtol=1e-12, atol=1e-14, equal_nan=True)


@pytest.mark.parametrize("k", [300, 101, 64, 25, 200, 333])
def test_null_ratio_fast_path_equals_exact_selection(eng, k):
    """The thread-per-median kernel (15-bit order-preserving codes, packed fp16 bisection) against the exact
    warp-cooperative 64-bit selection (WCX_NULL_WARP=1) and the oracle on data built to break it: heavy ties
    (quantised values -> many reference bins inside one code cell), a constant column, a column with NaN and one with
    inf, -1 fillers, rows whose indexes are all the same bin, odd / even / unaligned k."""
    rng = np.random.default_rng(k)
    n, s = 6000, 24
    per = np.array([1500, 1200, 900, 700, 500, 400] + [50] * 16)
    x = 1.0 + 0.05 * rng.standard_normal((n, s))
    x[:, 3] = np.round(x[:, 3] * 64) / 64          # ~20 distinct values: every median sits in a tie
    x[:, 5] = 0.75                                   # no spread
    x[rng.integers(0, n, 40), 7] = np.nan
    x[rng.integers(0, n, 5), 9] = np.inf
    x[:, 11] = np.round(x[:, 11] * 4096) / 4096     # mild ties
    cum = np.cumsum(per)
    rows = 700
    idx = rng.integers(0, n - 1500, size=(rows, k)).astype(np.int32)
    idx[5, :] = 17                                   # placeholder-like row
    idx[6, k // 2:] = -1                             # fillers wrap to the last bin
    idx[7, :] = np.arange(k)                         # consecutive bins
    ids = list(range(s))
    eng.load(x, per, cum)
    want = np_oracle.null_ratios(x, idx, 100, 100 + rows, ids)
    got = eng.null_ratios(100, 100 + rows, k, ids, idx=idx)
    os.environ["WCX_NULL_WARP"] = "1"
    try:
        legacy = eng.null_ratios(100, 100 + rows, k, ids, idx=idx)
    finally:
        del os.environ["WCX_NULL_WARP"]
    assert np.array_equal(got, legacy, equal_nan=True)  # same medians, same final division / log2
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-14, equal_nan=True)
    assert np.isnan(got[:, 7]).any() and not np.isnan(got[:, 0]).any()
