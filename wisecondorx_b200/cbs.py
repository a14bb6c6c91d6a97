"""Host side of the CUDA circular binary segmentation: the reference's `exec_cbs`
(predict_tools.py:242-275) with the R bridge (`exec_R`, overall_tools.py:65-80 -> include/CBS.R)
replaced by `wcx_cbs_segment`.  The pre-/post-processing of CBS.R (ratio == 0 -> NA, weight == 0 ->
1, all-NA chromosomes dropped, segments split over NA runs longer than int(2e6 / binsize), weighted
segment means, 0-based half-open coordinates) is restated here line by line."""
from __future__ import annotations

import ctypes
import logging
import math

import numpy as np

from . import _lib, predict_tools


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


# ---------------------------------------------------------------------------------------------
# DNAcopy's sequential stopping boundary: segment() hands `sbdry = getbdry(eta, nperm, max.ones)` (eta = 0.05,
# max.ones = floor(nperm * alpha) + 1) to the change-point finder.  Row j of the triangular table (nrejc = j - 1
# tolerated exceedances) holds the permutation counts b_0 <= ... <= b_{j-1}: with k exceedances seen, the test is
# declared significant once b_k permutations are done.  The b_k of a row share one per-boundary level eta0
# (b_k = first i with P(at most k of j exceedances among the first i of nperm) <= eta0, a hypergeometric tail) and eta0 is
# tuned until the probability of stopping early although j exceedances exist equals eta.  Restated from Venkatraman &
# Olshen 2007 with exact crossing probabilities; the R / Fortran sources are not available here (see oracle/cbs_oracle.py).
# ---------------------------------------------------------------------------------------------
_BOUNDARY_CACHE = {}


def _lchoose(n, k):
    if k < 0 or k > n:
        return -math.inf
    return math.lgamma(n + 1.0) - math.lgamma(k + 1.0) - math.lgamma(n - k + 1.0)


def _hyper_cdf(k, ones, nperm, i):
    den = _lchoose(nperm, i)
    return float(sum(math.exp(_lchoose(ones, x) + _lchoose(nperm - ones, i - x) - den) for x in range(min(k, ones, i) + 1)))


def _row_for_level(nperm, eta0, ones):
    row, start = [], 1
    for k in range(ones):
        if _hyper_cdf(k, ones, nperm, nperm) > eta0:
            row.append(nperm)
            start = nperm
            continue
        a, z = start, nperm
        while a < z:  # the tail probability falls with i
            mid = (a + z) // 2
            if _hyper_cdf(k, ones, nperm, mid) <= eta0:
                z = mid
            else:
                a = mid + 1
        row.append(a)
        start = min(nperm, a + 1)
    return row


def _early_stop_probability(nperm, ones, row):
    """P(some exceedance k + 1 of `ones` uniformly placed ones comes after permutation row[k])."""
    cum = np.ones(nperm + 1)
    log_scale = 0.0
    for k in range(ones):
        ways = np.zeros(nperm + 1)
        ways[1:row[k] + 1] = cum[0:row[k]]
        cum = np.cumsum(ways)
        if cum[-1] <= 0.0:
            return 1.0
        log_scale += math.log(cum[-1])
        cum /= cum[-1]
    return 1.0 - math.exp(log_scale - _lchoose(nperm, ones))


def sequential_boundary(eta, nperm, max_ones, tol=1e-2):
    key = (float(eta), int(nperm), int(max_ones), float(tol))
    if key not in _BOUNDARY_CACHE:
        table = [nperm - int(nperm * eta)]
        level = eta
        for ones in range(2, max_ones + 1):
            hi = level * 1.1
            p_hi = _early_stop_probability(nperm, ones, _row_for_level(nperm, hi, ones))
            lo = level * 0.25
            row = _row_for_level(nperm, lo, ones)
            p_lo = _early_stop_probability(nperm, ones, row)
            while (hi - lo) / lo > tol:
                level = lo + (hi - lo) * (eta - p_lo) / (p_hi - p_lo)
                row = _row_for_level(nperm, level, ones)
                p = _early_stop_probability(nperm, ones, row)
                if p > eta:
                    hi, p_hi = level, p
                else:
                    lo, p_lo = level, p
            table.extend(row)
        _BOUNDARY_CACHE[key] = np.ascontiguousarray(table, dtype=np.int32)
    return _BOUNDARY_CACHE[key]


def _segment_flat(y, w, off, ids, alpha, nperm, seed, ctx, sequential=True, eta=0.05):
    """One wcx_cbs_segment call over the series y[off[s]:off[s + 1]] (NA-free, weights w, permutation stream ids[s]).
    Returns (ends, nseg): the ascending exclusive segment ends of all series back to back and their number per series."""
    ctx = ctx or _lib.default_context(0)
    L = _lib.load()
    if sequential:
        table = sequential_boundary(eta, nperm, int(math.floor(nperm * alpha)) + 1)
        _lib.check(L.wcx_cbs_set_boundary(ctx.handle, _ptr(table), len(table)))
    else:
        _lib.check(L.wcx_cbs_set_boundary(ctx.handle, None, 0))
    ns = len(off) - 1
    total = int(off[-1])
    y = np.ascontiguousarray(y, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    off = np.ascontiguousarray(off, dtype=np.int64)
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    ends = np.zeros(max(total, 1), dtype=np.int32)
    nseg = np.zeros(max(ns, 1), dtype=np.int32)
    _lib.check(L.wcx_cbs_segment(ctx.handle, _ptr(y), _ptr(w), _ptr(off), ns, _ptr(ids), float(alpha), int(nperm),
                                 int(seed) & 0xFFFFFFFF, _ptr(ends), _ptr(nseg)))
    predict_tools.accumulate_ms(ctx, ("cbs",))
    return ends, nseg[:ns]


def segment_series(series, series_ids=None, alpha=1e-4, nperm=10000, seed=0, ctx: _lib.Context | None = None,
                   sequential: bool = True, eta: float = 0.05):
    """Segments a batch of NA-free (y, w) series on the GPU.  Returns a list of int32 arrays with
    the ascending exclusive segment ends of each series.  sequential: DNAcopy's early-stopping decision rule
    (default, as in segment()); False counts the exceedances over all nperm permutations."""
    ns = len(series)
    off = np.concatenate([[0], np.cumsum([len(y) for y, _ in series])]).astype(np.int64)
    y = np.concatenate([np.asarray(a, dtype=np.float64) for a, _ in series]) if ns else np.zeros(0)
    w = np.concatenate([np.asarray(b, dtype=np.float64) for _, b in series]) if ns else np.zeros(0)
    ends, nseg = _segment_flat(y, w, off, np.arange(ns) if series_ids is None else series_ids, alpha, nperm, seed, ctx,
                               sequential, eta)
    cut = np.concatenate([[0], np.cumsum(nseg)]).astype(np.int64)
    return [ends[cut[s]:cut[s + 1]].copy() for s in range(ns)]


def cbs_stats(ctx: _lib.Context | None = None):
    ctx = ctx or _lib.default_context(0)
    out = np.zeros(6, dtype=np.int64)
    _lib.check(_lib.load().wcx_cbs_stats(ctx.handle, _ptr(out)))
    return dict(zip(["rounds", "segments_tested", "perm_tests", "t_tests", "permutations", "launches"], out.tolist()))


def cbs_segments(results_r, results_w, ref_gender, alpha, binsize, seed=None, nperm=10000, ctx=None):
    """CBS.R as a function: [[chr (0-based), s, e (exclusive), r], ...]."""
    return cbs_segments_batch([(results_r, results_w, ref_gender)], alpha, binsize, seed, nperm, ctx)[0]


_noted = [False]


def _note_not_bit_compatible():
    """Once per process: the segmentation follows DNAcopy's algorithm but not R's random stream."""
    if not _noted[0]:
        logging.info("Segmentation runs on the GPU (circular binary segmentation as in DNAcopy::segment, hybrid p-value and "
                     "sequential stopping boundary included; the permutations come from a Philox stream, not R's Mersenne "
                     "Twister): breakpoints of borderline segments can differ from the reference's R run")
        _noted[0] = True


def cbs_segments_batch(samples, alpha, binsize, seed=None, nperm=10000, ctx=None):
    """CBS.R for a batch of samples [(results_r, results_w, ref_gender), ...] with ONE device call over all
    (sample, chromosome) series.  The permutation streams are keyed by (seed, chromosome), not by the position of a
    series in the batch, so every sample gets the segments it would get alone.

    CBS.R's own work around DNAcopy::segment runs on host threads in the library (csrc/host_cbs.cu): :30-63 (ratio == 0
    -> NA, weight == 0 -> 1, the NA-free series of every chromosome, chromosomes without data dropped) before the device
    call, :80-129 (segments cut at NA runs longer than int(2e6 / binsize), weighted means, 0-based half-open coordinates)
    after it."""
    import os
    _note_not_bit_compatible()
    seed_i = 0 if seed is None else int(seed)
    L = _lib.load()
    threads = max(1, min(16, len(os.sched_getaffinity(0))))
    n = len(samples)
    if n == 0:
        return []
    flat_r, flat_w, offs_s = [], [], []
    for results_r, results_w, ref_gender in samples:
        nchr = 24 if ref_gender == "M" else 23  # CBS.R:30-34
        if len(results_r) < nchr or len(results_w) < nchr:
            raise IndexError("list index out of range")
        offs_s.append(np.concatenate([[0], np.cumsum([len(x) for x in results_r[:nchr]])]).astype(np.int64))
        # (views of the rows the result assembly wrote: no copies)
        flat_r.append(np.ascontiguousarray(predict_tools.flatten(results_r[:nchr])))
        flat_w.append(np.ascontiguousarray(predict_tools.flatten(results_w[:nchr])))
        if len(flat_w[-1]) != len(flat_r[-1]):
            raise ValueError("results_r and results_w differ in length")
    nchr_s = np.array([len(o) - 1 for o in offs_s], dtype=np.int64)
    offs_all = np.ascontiguousarray(np.concatenate(offs_s))
    offs_at = np.concatenate([[0], np.cumsum(nchr_s + 1)]).astype(np.int64)
    slot_first = np.concatenate([[0], np.cumsum(nchr_s)]).astype(np.int64)  # first (sample, chromosome) slot of a sample
    r_ptrs = np.array([a.ctypes.data for a in flat_r], dtype=np.uintp)
    w_ptrs = np.array([a.ctypes.data for a in flat_w], dtype=np.uintp)
    na_thresh = int((binsize / 2000000.0) ** -1)  # CBS.R:95
    counts = np.zeros(int(slot_first[-1]), dtype=np.int64)
    _lib.check(L.wcx_cbs_pack_count(_ptr(r_ptrs), _ptr(offs_all), _ptr(offs_at), n, _ptr(counts), threads))
    at = np.concatenate([[0], np.cumsum(np.add.reduceat(counts, slot_first[:-1]))]).astype(np.int64)  # every sample has >= 23 slots
    total = int(at[-1])
    # the series of all samples back to back in two page-locked vectors (they cross PCIe in the device call)
    alloc = _lib.pinned.empty if n >= 8 else (lambda shape: np.empty(shape, dtype=np.float64))
    y, w = alloc((max(total, 1),)), alloc((max(total, 1),))
    pos = np.empty(max(total, 1), dtype=np.int32)
    gaps = np.zeros(len(counts), dtype=np.int64)
    _lib.check(L.wcx_cbs_pack(_ptr(r_ptrs), _ptr(w_ptrs), _ptr(offs_all), _ptr(offs_at), n, _ptr(at), na_thresh, _ptr(y), _ptr(w),
                              _ptr(pos), _ptr(gaps), threads))
    series = np.flatnonzero(counts > 0)  # chromosomes without data are dropped (CBS.R:56-63)
    off = np.concatenate([[0], np.cumsum(counts[series])]).astype(np.int64)
    chrom = np.concatenate([np.arange(k) for k in nchr_s])[series]
    sample_of = np.repeat(np.arange(n), nchr_s)[series]
    chr_start = np.ascontiguousarray(np.concatenate([o[:-1] for o in offs_s])[series], dtype=np.int64)
    ends, nseg = _segment_flat(y[:total], w[:total], off, chrom, alpha, nperm, seed_i, ctx)
    ends, nseg = np.ascontiguousarray(ends, dtype=np.int32), np.ascontiguousarray(nseg, dtype=np.int32)
    if len(series) and int(nseg.min()) < 1:
        raise _lib.WcxError("wcx_cbs_segment returned a series without segments")
    slot = np.concatenate([[0], np.cumsum(nseg.astype(np.int64) + gaps[series])]).astype(np.int64)
    cap = max(int(slot[-1]), 1)
    out_series = np.full(cap, -1, dtype=np.int32)
    out_s, out_e, out_r = np.empty(cap, dtype=np.int64), np.empty(cap, dtype=np.int64), np.empty(cap, dtype=np.float64)
    _lib.check(L.wcx_cbs_unpack(_ptr(pos), _ptr(y), _ptr(w), _ptr(off), len(series), _ptr(ends), _ptr(nseg), _ptr(chr_start), _ptr(slot),
                                na_thresh, _ptr(out_series), _ptr(out_s), _ptr(out_e), _ptr(out_r), threads))
    used = np.flatnonzero(out_series >= 0)
    ser = out_series[used]
    rows = [list(t) for t in zip(chrom[ser].tolist(), out_s[used].tolist(), out_e[used].tolist(), out_r[used].tolist())]
    cut = np.concatenate([[0], np.cumsum(np.bincount(sample_of[ser], minlength=n))]).astype(int)
    return [rows[cut[j]:cut[j + 1]] for j in range(n)]  # [[chr (0-based), s, e (exclusive), r], ...] per sample


def exec_cbs(rem_input, results, engine: predict_tools.PredictEngine | None = None, nperm=10000):
    """Drop-in for predict_tools.exec_cbs (reference predict_tools.py:242-263): segments
    results["results_r"] with weights results["results_w"], then attaches the between-sample
    segment z-scores -> [[chr, s, e, z, r], ...]."""
    args = rem_input["args"]
    results_c = cbs_segments(results["results_r"], results["results_w"], str(rem_input["ref_gender"]), float(args.alpha),
                             float(rem_input["binsize"]), getattr(args, "seed", None), nperm,
                             engine.ctx if engine else None)
    segment_z = predict_tools.get_z_score(results_c, results, engine)
    return [results_c[i][:3] + [segment_z[i]] + [results_c[i][3]] for i in range(len(results_c))]


def exec_cbs_batch(rem_inputs, results_list, engine: predict_tools.PredictEngine | None = None, nperm=10000):
    """exec_cbs for a batch of samples that share args (alpha, seed) and the reference: one CBS call, then the
    segment z-scores per sample."""
    if not rem_inputs:
        return []
    args = rem_inputs[0]["args"]
    segs = cbs_segments_batch([(res["results_r"], res["results_w"], str(rem["ref_gender"])) for rem, res in zip(rem_inputs, results_list)],
                              float(args.alpha), float(rem_inputs[0]["binsize"]), getattr(args, "seed", None), nperm,
                              engine.ctx if engine else None)
    out = [None] * len(segs)
    # samples of one reference gender share their null-ratio array (resident on the device): one z-score call per gender,
    # in chunks that stay below the kernel's segment limit
    groups = {}
    for j in range(len(segs)):
        groups.setdefault(id(results_list[j]["results_nr"]["dense"]), []).append(j)
    for js in groups.values():
        chunk, nseg = [], 0
        for j in js + [None]:
            if j is None or nseg + len(segs[j]) > 60000:
                for jj, z in zip(chunk, predict_tools.get_z_score_batch([(segs[i], results_list[i]) for i in chunk], engine)):
                    out[jj] = [segs[jj][i][:3] + [z[i]] + [segs[jj][i][3]] for i in range(len(segs[jj]))]
                chunk, nseg = [], 0
            if j is not None:
                chunk.append(j)
                nseg += len(segs[j])
    return out
