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


def segment_series(series, series_ids=None, alpha=1e-4, nperm=10000, seed=0, ctx: _lib.Context | None = None,
                   sequential: bool = True, eta: float = 0.05):
    """Segments a batch of NA-free (y, w) series on the GPU.  Returns a list of int32 arrays with
    the ascending exclusive segment ends of each series.  sequential: DNAcopy's early-stopping decision rule
    (default, as in segment()); False counts the exceedances over all nperm permutations."""
    ctx = ctx or _lib.default_context(0)
    L = _lib.load()
    if sequential:
        table = sequential_boundary(eta, nperm, int(math.floor(nperm * alpha)) + 1)
        _lib.check(L.wcx_cbs_set_boundary(ctx.handle, _ptr(table), len(table)))
    else:
        _lib.check(L.wcx_cbs_set_boundary(ctx.handle, None, 0))
    ns = len(series)
    lens = [len(y) for y, _ in series]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    total = int(off[-1])
    y = np.ascontiguousarray(np.concatenate([np.asarray(a, dtype=np.float64) for a, _ in series]) if ns else np.zeros(0))
    w = np.ascontiguousarray(np.concatenate([np.asarray(b, dtype=np.float64) for _, b in series]) if ns else np.zeros(0))
    ids = np.ascontiguousarray(np.arange(ns) if series_ids is None else series_ids, dtype=np.int32)
    ends = np.zeros(max(total, 1), dtype=np.int32)
    nseg = np.zeros(max(ns, 1), dtype=np.int32)
    _lib.check(L.wcx_cbs_segment(ctx.handle, _ptr(y), _ptr(w), _ptr(off), ns, _ptr(ids), float(alpha), int(nperm),
                                 int(seed) & 0xFFFFFFFF, _ptr(ends), _ptr(nseg)))
    predict_tools.accumulate_ms(ctx, ("cbs",))
    out, o = [], 0
    for s in range(ns):
        out.append(ends[o:o + nseg[s]].copy())
        o += int(nseg[s])
    return out


def cbs_stats(ctx: _lib.Context | None = None):
    ctx = ctx or _lib.default_context(0)
    out = np.zeros(6, dtype=np.int64)
    _lib.check(_lib.load().wcx_cbs_stats(ctx.handle, _ptr(out)))
    return dict(zip(["rounds", "segments_tested", "perm_tests", "t_tests", "permutations", "launches"], out.tolist()))


def _cbs_prepare(results_r, results_w, ref_gender):
    """CBS.R:30-63 for one sample: per chromosome the ratio / weight vectors, the NA mask and the NA-free series."""
    nchr = 24 if ref_gender == "M" else 23  # CBS.R:30-34
    prepared, series, ids = [], [], []
    for c in range(nchr):
        ratio = np.asarray(results_r[c], dtype=np.float64)
        wts = np.asarray(results_w[c], dtype=np.float64).copy()
        na = ratio == 0  # CBS.R:41
        wts[wts == 0] = 1.0  # CBS.R:42 -- 1^-99 is 1 in R
        if na.all():  # CBS.R:56-63
            continue
        keep = np.flatnonzero(~na)
        prepared.append((c, ratio, wts, na, keep))
        series.append((ratio[keep], wts[keep]))
        ids.append(c)
    return prepared, series, ids


def _cbs_finish(prepared, all_ends, binsize):
    """CBS.R:80-129: split the segments over long NA runs, weighted segment means, 0-based half-open coordinates.
    The NA runs of a chromosome are located once (CBS.R does it per segment, :86-101, with the same result: a segment
    starts and ends on a non-NA bin, so a run lies inside it or outside)."""
    na_thresh = int((binsize / 2000000.0) ** -1)  # CBS.R:95
    out = []
    for (c, ratio, wts, na, keep), ends in zip(prepared, all_ends):
        d = np.diff(na.astype(np.int8))
        run_first = np.flatnonzero(d == 1) + 1   # first NA bin of a run (0-based) = CBS.R's start.pos (1-based last bin before it)
        run_after = np.flatnonzero(d == -1) + 1  # first bin after a run (0-based)   = CBS.R's end.pos (1-based last NA bin)
        if len(na) and na[0]:
            run_after = run_after[1:]  # a run at the chromosome start has no beginning inside any segment
        if len(na) and na[-1]:
            run_first = run_first[:-1]
        long_run = (run_after - run_first) > na_thresh  # CBS.R:95
        run_first, run_after = run_first[long_run], run_after[long_run]
        starts = np.concatenate([[0], ends[:-1]]).astype(np.int64)
        for a, b in zip(starts.tolist(), np.asarray(ends).tolist()):
            start_i, end_i = int(keep[a]) + 1, int(keep[b - 1]) + 1  # DNAcopy loc.start / loc.end (1-based)
            lo = int(np.searchsorted(run_first, start_i - 1, "right")) if len(run_first) else 0
            hi = int(np.searchsorted(run_first, end_i - 1, "left")) if len(run_first) else 0
            if hi > lo:
                inv_start = [start_i] + run_after[lo:hi].tolist()  # CBS.R:100-101
                inv_end = run_first[lo:hi].tolist() + [end_i]
            else:
                inv_start, inv_end = [start_i], [end_i]
            for s1, e1 in zip(inv_start, inv_end):
                if e1 - s1 <= 0:  # CBS.R:103
                    continue
                yy, ww = ratio[s1 - 1:e1], wts[s1 - 1:e1]
                m = yy != 0
                r = float(np.sum(yy[m] * ww[m]) / np.sum(ww[m])) if m.any() else float("nan")  # CBS.R:122-127
                out.append([c, int(s1) - 1, int(e1), r])  # CBS.R:129, predict_tools.py:266-275
    return out


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
    series in the batch, so every sample gets the segments it would get alone."""
    from .predict_control import _map_threads
    _note_not_bit_compatible()
    seed_i = 0 if seed is None else int(seed)
    preps, series, ids, counts = [], [], [], []
    for p, s, i in _map_threads(lambda t: _cbs_prepare(*t), samples):
        preps.append(p); series += s; ids += i; counts.append(len(s))
    all_ends = segment_series(series, ids, alpha, nperm, seed_i, ctx)
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(int)
    return _map_threads(lambda j: _cbs_finish(preps[j], all_ends[offs[j]:offs[j + 1]], binsize), range(len(preps)))


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
