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


class _Prepared:
    """CBS.R:30-63 for one sample, on the concatenated bin axis of its chromosomes: `na` (ratio == 0), `cols` (positions
    of the other bins), `y` / `w` (their ratios and weights, weight 0 -> 1), `base[c]` = first entry of chromosome c in
    them, `ids` = the chromosomes that are not all NA (the others are dropped, CBS.R:56-63)."""
    __slots__ = ("offs", "na", "cols", "base", "y", "w", "ids")


def _cbs_prepare_flat(r_flat, w_flat, offs, gather=True):
    p = _Prepared()
    p.offs = offs
    p.na = r_flat == 0  # CBS.R:41
    p.cols = np.flatnonzero(~p.na)
    p.base = np.searchsorted(p.cols, offs)
    p.ids = np.flatnonzero(np.diff(p.base) > 0)
    if gather:
        _cbs_gather(p, r_flat, w_flat, np.empty(len(p.cols)), np.empty(len(p.cols)))
    return p


def _cbs_gather(p, r_flat, w_flat, y_out, w_out):
    """The NA-free ratios and weights of a prepared sample, written where the caller wants them (a batch lines the
    samples up in the two vectors of its one device call)."""
    np.take(r_flat, p.cols, out=y_out, mode="clip")
    np.take(w_flat, p.cols, out=w_out, mode="clip")
    w_out[w_out == 0] = 1.0  # CBS.R:42 -- 1^-99 is 1 in R
    p.y, p.w = y_out, w_out


def _cbs_prepare(results_r, results_w, ref_gender):
    """CBS.R:30-63 for one sample given as per-chromosome lists.  Returns (prepared, series, ids): the NA-free
    (ratio, weight) series of the chromosomes `ids` that have any data."""
    nchr = 24 if ref_gender == "M" else 23  # CBS.R:30-34
    if len(results_r) < nchr or len(results_w) < nchr:
        raise IndexError("list index out of range")
    r_flat = predict_tools.flatten(results_r[:nchr])
    w_flat = predict_tools.flatten(results_w[:nchr])
    offs = np.concatenate([[0], np.cumsum([len(x) for x in results_r[:nchr]])]).astype(np.int64)
    p = _cbs_prepare_flat(r_flat, w_flat, offs)
    series = [(p.y[p.base[c]:p.base[c + 1]], p.w[p.base[c]:p.base[c + 1]]) for c in p.ids]
    return p, series, [int(c) for c in p.ids]


def _cbs_finish(p, all_ends, binsize):
    """CBS.R:80-129: split the segments over long NA runs, weighted segment means, 0-based half-open coordinates.
    all_ends[i] = ascending exclusive ends of the segments of chromosome p.ids[i] in its NA-free series.

    The NA runs of the whole sample are located once (CBS.R does it per segment, :86-101, with the same result: a
    segment starts and ends on a non-NA bin, so a run lies inside it or outside).  CBS.R only sees runs that begin and
    end inside the segment, i.e. inside the chromosome: a run that touches a chromosome boundary is no run."""
    na_thresh = int((binsize / 2000000.0) ** -1)  # CBS.R:95
    offs, cols, base = p.offs, p.cols, p.base
    d = np.diff(p.na.view(np.int8))
    first = np.flatnonzero(d == 1) + 1   # first NA bin of a run           (CBS.R's start.pos, 1-based: the bin before it)
    after = np.flatnonzero(d == -1) + 1  # first bin after the run, 0-based (CBS.R's end.pos, 1-based: the last NA bin)
    if len(p.na) and p.na[0]:
        after = after[1:]
    if len(p.na) and p.na[-1]:
        first = first[:-1]
    sel = (after - first) > na_thresh  # CBS.R:95
    first, after = first[sel], after[sel]
    if len(first):  # no chromosome start inside [first, after]: the run begins and ends within one chromosome
        inside = np.searchsorted(offs, first, "left") == np.searchsorted(offs, after, "right")
        first, after = first[inside], after[inside]
    # segments of all chromosomes: entries [a, b) of the NA-free vectors, first / last bin on the concatenated axis
    if not len(p.ids):
        return []
    counts = np.array([len(e) for e in all_ends], dtype=np.int64)
    ends = np.concatenate([np.asarray(e, dtype=np.int64) for e in all_ends])
    if not len(ends):
        return []
    chrom = np.repeat(np.asarray(p.ids, dtype=np.int64), counts)
    seg_a = np.empty(len(ends), dtype=np.int64)
    seg_a[1:] = ends[:-1]
    seg_a[(np.cumsum(counts) - counts)[counts > 0]] = 0  # the first segment of a chromosome starts at its first entry
    seg_a += base[chrom]
    seg_b = ends + base[chrom]
    gs, ge = cols[seg_a], cols[seg_b - 1]  # DNAcopy loc.start / loc.end (here 0-based, concatenated axis)
    lo = np.searchsorted(first, gs, "right") if len(first) else np.zeros(len(gs), dtype=np.int64)
    hi = np.searchsorted(first, ge, "left") if len(first) else lo
    yw = p.y * p.w
    out = []
    for c, a, b, s, e, l, h in zip(chrom.tolist(), seg_a.tolist(), seg_b.tolist(), gs.tolist(), ge.tolist(), lo.tolist(), hi.tolist()):
        a0 = int(offs[c])
        if h <= l:
            if e - s <= 0:  # CBS.R:103
                continue
            # the non-zero ratios of bins [s, e] are the entries [a, b) of the NA-free vectors (CBS.R:122-127)
            out.append([c, s - a0, e - a0 + 1, float(np.sum(yw[a:b]) / np.sum(p.w[a:b]))])  # CBS.R:129, predict_tools.py:266-275
            continue
        inv_start = [s + 1] + after[l:h].tolist()  # CBS.R:100-101 (1-based on the concatenated axis)
        inv_end = first[l:h].tolist() + [e + 1]
        for s1, e1 in zip(inv_start, inv_end):
            if e1 - s1 <= 0:  # CBS.R:103
                continue
            a1, b1 = np.searchsorted(cols, [s1 - 1, e1])
            r = float(np.sum(yw[a1:b1]) / np.sum(p.w[a1:b1])) if b1 > a1 else float("nan")
            out.append([c, s1 - 1 - a0, e1 - a0, r])
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
    _map_threads = predict_tools.map_threads
    _note_not_bit_compatible()
    seed_i = 0 if seed is None else int(seed)

    def prepare(t):
        results_r, results_w, ref_gender = t
        nchr = 24 if ref_gender == "M" else 23  # CBS.R:30-34
        if len(results_r) < nchr or len(results_w) < nchr:
            raise IndexError("list index out of range")
        offs = np.concatenate([[0], np.cumsum([len(x) for x in results_r[:nchr]])]).astype(np.int64)
        flat = predict_tools.flatten(results_r[:nchr]), predict_tools.flatten(results_w[:nchr])
        return _cbs_prepare_flat(flat[0], flat[1], offs, gather=False), flat

    prepared = _map_threads(prepare, samples)
    preps = [t[0] for t in prepared]
    lens = [np.diff(p.base)[p.ids] for p in preps]
    off = np.concatenate([[0], np.cumsum(np.concatenate(lens))]).astype(np.int64) if preps else np.zeros(1, dtype=np.int64)
    ids = np.concatenate([p.ids for p in preps]) if preps else np.zeros(0, dtype=np.int32)
    # the series of all samples back to back in two page-locked vectors, every sample gathered into its place
    alloc = _lib.pinned.empty if len(preps) >= 8 else (lambda shape: np.empty(shape, dtype=np.float64))
    y, w = alloc((max(int(off[-1]), 1),)), alloc((max(int(off[-1]), 1),))
    at = np.concatenate([[0], np.cumsum([len(p.cols) for p in preps])]).astype(np.int64)
    _map_threads(lambda j: _cbs_gather(preps[j], prepared[j][1][0], prepared[j][1][1], y[at[j]:at[j + 1]], w[at[j]:at[j + 1]]),
                 range(len(preps)))
    ends, nseg = _segment_flat(y, w, off, ids, alpha, nperm, seed_i, ctx)
    cut = np.concatenate([[0], np.cumsum(nseg)]).astype(np.int64)  # first segment end of every series
    first = np.concatenate([[0], np.cumsum([len(p.ids) for p in preps])]).astype(np.int64)  # first series of every sample

    def finish(j):
        return _cbs_finish(preps[j], [ends[cut[s]:cut[s + 1]] for s in range(first[j], first[j + 1])], binsize)

    return _map_threads(finish, range(len(preps)))


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
