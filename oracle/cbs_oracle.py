"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the CBS path of WisecondorX `predict`.

    *** PARITY UNPINNED ***

The reference runs circular binary segmentation by shelling out to R:
``exec_cbs`` (predict_tools.py:242-263) -> ``exec_R`` (overall_tools.py:65-80) -> ``include/CBS.R``
-> ``DNAcopy::segment(CNA(...), alpha=alpha, verbose=1, weights=w)`` (CBS.R:70-73).  DNAcopy is a
third-party Bioconductor package (pinned ``bioconductor-dnacopy ==1.76``, conda.yml:14); neither
its source nor R is available in the build container and the reference ships no test, fixture or
golden vector for this boundary.  This file therefore restates

  (1) the reference's own pre-/post-processing (CBS.R:21-132), line by line, and
  (2) the published algorithm behind ``segment`` -- Olshen et al., Biostatistics 2004 (CBS) and
      Venkatraman & Olshen, Bioinformatics 2007 (hybrid p-value) -- with the documented defaults
      of DNAcopy 1.76 (nperm=10000, p.method="hybrid", min.width=2, kmax=25, nmin=200, eta=0.05,
      trim=0.025, undo.splits="none"; only alpha and weights are overridden by CBS.R:73),

and is anchored on the reference's call sites and the soft known answer in
docs/include/example.bed (tests/test_cbs_*.py).  It is the oracle the CUDA CBS is compared with
bit for bit (same counter-based RNG, same arithmetic order).

Deliberate, documented differences from DNAcopy:
  * R's Mersenne-Twister stream cannot be reproduced; permutations use Philox4x32-10 keyed by
    (seed, segment start, segment end, test id, permutation index).
  * DNAcopy's sequential boundary (``getbdry``, eta = 0.05: a test is declared significant as soon as the
    number of permutations seen without the (j + 1)-th exceedance reaches ``sbdry``) is restated from the
    description in Venkatraman & Olshen 2007 and the structure of ``segment()`` (``max.ones = floor(nperm *
    alpha) + 1``, triangular table indexed by ``nrejc * (nrejc + 1) / 2 + nrej``) with EXACT hypergeometric
    crossing probabilities (`seq_boundary`); DNAcopy approximates them beyond three exceedances, so for
    alpha >= 3 / nperm a boundary value may differ by a permutation or two.  ``sequential=False`` gives the
    plain count ``#exceedances <= nrejc`` over all nperm permutations.
  * the maximal statistic is found by brute force over all arcs (DNAcopy uses a block algorithm
    with the same result); ties resolve to the smallest (start, end).
All sums are sequential left-to-right (np.cumsum), no pairwise summation, so the CUDA kernels can
reproduce every rounding.
"""
from __future__ import annotations

import math

import numpy as np

# ---------------------------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al. 2011) -- identical integer arithmetic on host and device
# ---------------------------------------------------------------------------------------------
_M0, _M1 = 0xD2511F53, 0xCD9E8D57
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF


def philox4x32(counter, key):
    c0, c1, c2, c3 = counter
    k0, k1 = key
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> 32, p0 & _MASK
        hi1, lo1 = p1 >> 32, p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & _MASK, lo1, (hi0 ^ c3 ^ k1) & _MASK, lo0
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return c0, c1, c2, c3


class PermStream:
    """Random 32-bit words of one permutation: word t = philox(counter=(t // 4, perm, lo, hi),
    key=(seed, test))[t % 4]."""

    def __init__(self, seed, test, lo, hi, perm):
        self.key = (seed & _MASK, test & _MASK)
        self.base = (perm & _MASK, lo & _MASK, hi & _MASK)
        self.t = 0
        self.buf = None

    def next_u32(self):
        if self.t % 4 == 0:
            self.buf = philox4x32((self.t // 4, *self.base), self.key)
        v = self.buf[self.t % 4]
        self.t += 1
        return v

    def below(self, i):
        """uniform integer in [0, i): (u32 * i) >> 32"""
        return (self.next_u32() * i) >> 32


# ---------------------------------------------------------------------------------------------
# tail probability of the CBS statistic (Siegmund 1988 / Yao 1989), DNAcopy `tailp`, `nu`, `it1tsq`
# ---------------------------------------------------------------------------------------------
def _pnorm(x):
    return 0.5 * math.erfc(-x / math.sqrt(2.0))


def nu(x, tol):
    if x > 0.01:
        lnu1 = math.log(2.0) - 2.0 * math.log(x)
        lnu0 = lnu1
        k = 2
        dk = 0.0
        for _ in range(k):
            dk += 1.0
            lnu1 -= 2.0 * _pnorm(-x * math.sqrt(dk) / 2.0) / dk
        while abs((lnu1 - lnu0) / lnu1) > tol:
            lnu0 = lnu1
            for _ in range(k):
                dk += 1.0
                lnu1 -= 2.0 * _pnorm(-x * math.sqrt(dk) / 2.0) / dk
            k *= 2
    else:
        lnu1 = -0.583 * x
    return math.exp(lnu1)


def _it1tsq(x, a):
    """integral of 1 / (t (1 - t))^2 from x to x + a"""
    y = x + a - 0.5
    v = (8.0 * y) / (1.0 - 4.0 * y * y) + 2.0 * math.log((1.0 + 2.0 * y) / (1.0 - 2.0 * y))
    y = x - 0.5
    return v - (8.0 * y) / (1.0 - 4.0 * y * y) - 2.0 * math.log((1.0 + 2.0 * y) / (1.0 - 2.0 * y))


def tailp(b, delta, m, ngrid=100, tol=1e-6):
    dincr = (0.5 - delta) / ngrid
    bsqrtm = b / math.sqrt(m)
    tl = 0.5 - dincr
    t = 0.5 - 0.5 * dincr
    acc = 0.0
    for _ in range(ngrid):
        tl += dincr
        t += dincr
        x = bsqrtm / math.sqrt(t * (1.0 - t))
        nux = nu(x, tol)
        acc += (nux * nux) * _it1tsq(tl, dincr)
    return 9.973557e-2 * (b * b * b) * math.exp(-b * b / 2.0) * acc


# ---------------------------------------------------------------------------------------------
# statistics
# ---------------------------------------------------------------------------------------------
def seq_sum(a):
    return float(np.cumsum(np.asarray(a, dtype=np.float64))[-1]) if len(a) else 0.0


def max_arc(sx, cw, n, al0, max_width=None):
    """max over arcs 0 <= i < j <= n, al0 <= j - i <= n - al0 (and j - i <= max_width, or
    >= n - max_width: the circular complement) of
    bss = (sx[j] - sx[i])^2 / (dw * (cw[n] - dw)), dw = cw[j] - cw[i].
    sx, cw have n + 1 entries (leading 0).  Ties: smallest i, then smallest j."""
    best, bi, bj = -1.0, 0, 0
    cwn = cw[n]
    for i in range(0, n):
        jlo, jhi = i + al0, min(n, i + n - al0)
        if jhi < jlo:
            continue
        j = np.arange(jlo, jhi + 1)
        if max_width is not None:
            wdt = j - i
            j = j[(wdt <= max_width) | (wdt >= n - max_width)]
            if len(j) == 0:
                continue
        s = sx[j] - sx[i]
        dw = cw[j] - cw[i]
        bss = (s * s) / (dw * (cwn - dw))
        k = int(np.argmax(bss))
        if bss[k] > best:
            best, bi, bj = float(bss[k]), i, int(j[k])
    return best, bi, bj


def _tstat(bss, tss, n):
    return bss / ((tss - bss) / (n - 2.0))


def perm_values(y, rw, stream):
    """wxperm: Fisher-Yates on y = x * sqrt(w) from the top, then px[i] = y_perm[i] / rw[i]."""
    n = len(y)
    py = y.copy()
    px = np.empty(n)
    for i in range(n - 1, -1, -1):
        j = stream.below(i + 1)
        tmp = py[i]
        py[i] = py[j]
        py[j] = tmp
        px[i] = py[i] / rw[i]
    return px


def perm_stat(px, ws, cw, tss_y, n, al0, max_width, tot_w, rtw):
    """t^2 statistic of permuted data: re-centre with the weighted mean of the permuted values."""
    sxp = np.concatenate([[0.0], np.cumsum(ws * px)])
    xbar = sxp[n] / tot_w
    sxc = sxp - xbar * (cw * rtw)  # cumulative weight up to t = cw[t] * sqrt(sum ws)
    tss = tss_y - tot_w * xbar * xbar
    bss, _, _ = max_arc(sxc, cw, n, al0, max_width)
    return _tstat(bss, tss, n)


def t_perm_p(x, ws, rw, n1, n2, alpha_nperm, seed, test, lo, hi):
    """Permutation t-test that the first n1 points of x differ from the remaining n2
    (DNAcopy `wtpermp`, restated for weights).  Returns the p-value."""
    nperm = alpha_nperm
    n = n1 + n2
    if n1 == 1 or n2 == 1:
        return 1.0
    wx = ws * x
    w1, w2 = seq_sum(ws[:n1]), seq_sum(ws[n1:])
    s1, s2 = seq_sum(wx[:n1]), seq_sum(wx[n1:])
    wt = w1 + w2
    xbar = (s1 + s2) / wt
    tss = seq_sum(ws * x * x) - wt * xbar * xbar
    if n1 <= n2:
        m1, pos0, wp = n1, 0, w1
        ostat = 0.99999 * abs(s1 / w1 - xbar)
        tstat = (ostat * ostat) * w1 * wt / w2
    else:
        m1, pos0, wp = n2, n1, w2
        ostat = 0.99999 * abs(s2 / w2 - xbar)
        tstat = (ostat * ostat) * w2 * wt / w1
    tstat = tstat / ((tss - tstat) / (n - 2.0))
    if tstat > 25.0 and m1 >= 10:
        return 0.0
    y = x * rw
    nrej = 0
    for p in range(nperm):
        st = PermStream(seed, test, lo, hi, p)
        py = y.copy()
        acc = 0.0
        # partial Fisher-Yates: draw the values landing on the m1 positions pos0 .. pos0 + m1 - 1
        for t in range(m1):
            i = n - 1 - t
            j = st.below(i + 1)
            tmp = py[i]
            py[i] = py[j]
            py[j] = tmp
            acc += rw[pos0 + t] * py[i]
        pstat = abs(acc / wp - xbar)
        if ostat <= pstat:
            nrej += 1
    return nrej / float(nperm)


# ---------------------------------------------------------------------------------------------
# sequential stopping boundary of the permutation test (DNAcopy getbdry / etabdry / pexceed)
# ---------------------------------------------------------------------------------------------
def _log_choose(n, k):
    if k < 0 or k > n:
        return -math.inf
    return math.lgamma(n + 1.0) - math.lgamma(k + 1.0) - math.lgamma(n - k + 1.0)


def _phyper_le(k, m, nperm, i):
    """P(at most k of the m exceedances fall among the first i of nperm permutations)."""
    den = _log_choose(nperm, i)
    return float(sum(math.exp(_log_choose(m, x) + _log_choose(nperm - m, i - x) - den) for x in range(0, min(k, m, i) + 1)))


def _eta_boundary(nperm, eta0, m):
    """etabdry: b[k] = first i with P(<= k exceedances among the first i | m in total) <= eta0, k = 0 .. m - 1."""
    b, lo = [], 1
    for k in range(m):
        a, z = lo, nperm  # P is non-increasing in i: bisection for the first i <= eta0
        if _phyper_le(k, m, nperm, z) > eta0:
            b.append(nperm)
            lo = nperm
            continue
        while a < z:
            mid = (a + z) // 2
            if _phyper_le(k, m, nperm, mid) <= eta0:
                z = mid
            else:
                a = mid + 1
        b.append(a)
        lo = min(nperm, a + 1)  # the Fortran loop advances i once per boundary
    return b


def _p_exceed(nperm, m, b):
    """pexceed: probability that m exceedances placed uniformly among nperm permutations let the test stop early,
    i.e. that for some k the (k + 1)-th exceedance comes after permutation b[k].  Exact (counting the placements
    t_1 < ... < t_m with t_k <= b[k] for all k)."""
    ways = np.zeros(nperm + 1)  # ways[t]: placements of the first k exceedances with t_k = t (scaled)
    cum = np.ones(nperm + 1)    # k = 0: one empty placement "ending" at every t >= 0
    log_scale = 0.0
    for k in range(m):
        ways[:] = 0.0
        ways[1:b[k] + 1] = cum[0:b[k]]
        cum = np.cumsum(ways)
        top = cum[-1]
        if top <= 0.0:
            return 1.0
        cum /= top
        log_scale += math.log(top)
    return 1.0 - math.exp(log_scale - _log_choose(nperm, m))


_SEQ_BOUNDARY = {}


def seq_boundary(eta, nperm, max_ones, tol=1e-2):
    """getbdry: triangular table; row j (1-based, j exceedances allowed to decide 'not significant', i.e.
    nrejc = j - 1) starts at j (j - 1) / 2 and holds j permutation counts."""
    key = (eta, nperm, max_ones, tol)
    if key in _SEQ_BOUNDARY:
        return _SEQ_BOUNDARY[key]
    out = [nperm - int(nperm * eta)]
    eta0 = eta
    for j in range(2, max_ones + 1):
        etahi = eta0 * 1.1
        bh = _eta_boundary(nperm, etahi, j)
        phi = _p_exceed(nperm, j, bh)
        etalo = eta0 * 0.25
        bl = _eta_boundary(nperm, etalo, j)
        plo = _p_exceed(nperm, j, bl)
        b = bl
        while (etahi - etalo) / etalo > tol:
            eta0 = etalo + (etahi - etalo) * (eta - plo) / (phi - plo)
            b = _eta_boundary(nperm, eta0, j)
            pe = _p_exceed(nperm, j, b)
            if pe > eta:
                etahi, phi = eta0, pe
            else:
                etalo, plo = eta0, pe
        out.extend(b)
    _SEQ_BOUNDARY[key] = out
    return out


def find_cpt(xs, ws, alpha, nperm, kmax, nmin, min_width, seed, lo, hi, stats=None, sbdry=None):
    """One call of DNAcopy's weighted change-point finder on a segment -> list of change points
    (offsets inside the segment)."""
    n = len(xs)
    if n < 2 * min_width:
        return []
    if abs(float(np.max(xs) - np.min(xs))) < 1.5e-8:  # isTRUE(all.equal(diff(range(x)), 0))
        return []
    tot_w = seq_sum(ws)
    avg = seq_sum(xs * ws) / tot_w
    x = xs - avg
    tss = seq_sum(ws * x * x)
    rw = np.sqrt(ws)
    rtw = math.sqrt(tot_w)
    cw = np.concatenate([[0.0], np.cumsum(ws) / rtw])
    sx = np.concatenate([[0.0], np.cumsum(ws * x)])
    bss, i1, i2 = max_arc(sx, cw, n, min_width)
    ostat = _tstat(bss, tss, n)
    ostat1 = math.sqrt(ostat) if ostat > 0 else 0.0
    ostat *= 0.99999
    if stats is not None:
        stats.append((lo, hi, ostat1, i1, i2))
    if ostat1 <= 0.1:
        return []
    width = i2 - i1
    l = min(width, n - width)
    accept = ostat1 >= 7.0 and l >= 10
    if not accept:
        hybrid = n > nmin
        y = x * rw
        tss_y = seq_sum(y * y)
        if hybrid:
            delta = (kmax + 1.0) / n
            pval1 = tailp(ostat1, delta, n)
            if pval1 > alpha:
                return []
            nrejc = int((alpha - pval1) * float(nperm))
            mw = kmax
        else:
            nrejc = int(alpha * float(nperm))
            mw = None
        nrej = 0
        k0 = nrejc * (nrejc + 1) // 2  # row of the boundary table for this rejection count
        use_bdry = sbdry is not None and k0 + nrejc < len(sbdry)
        for p in range(nperm):
            px = perm_values(y, rw, PermStream(seed, 0, lo, hi, p))
            if ostat <= perm_stat(px, ws, cw, tss_y, n, min_width, mw, tot_w, rtw):
                nrej += 1
                if nrej > nrejc:
                    return []
            if use_bdry and p + 1 >= sbdry[k0 + nrej]:
                break  # significant: the remaining permutations cannot change the decision but with probability eta
    if i2 == n:
        return [i1]
    if i1 == 0:
        return [i2]
    out = []
    if t_perm_p(x[:i2], ws[:i2], rw[:i2], i1, i2 - i1, nperm, seed, 1, lo, hi) <= alpha:
        out.append(i1)
    if t_perm_p(x[i1:], ws[i1:], rw[i1:], i2 - i1, n - i2, nperm, seed, 2, lo, hi) <= alpha:
        out.append(i2)
    return out


def segment_chromosome(y, w, alpha=1e-4, nperm=10000, kmax=25, nmin=200, min_width=2, seed=0, chrom=0,
                       stats=None, sequential=True, eta=0.05):
    """DNAcopy `changepoints` on the non-NA values of one chromosome -> segment end positions
    (exclusive, in the compacted index space) in ascending order."""
    y = np.asarray(y, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    n = len(y)
    ends = []
    stack = [(0, n)]
    s = (seed * 1000003 + chrom) & _MASK
    sbdry = seq_boundary(eta, nperm, int(math.floor(nperm * alpha)) + 1) if sequential else None  # segment(): max.ones
    while stack:
        lo, hi = stack.pop()
        cpts = find_cpt(y[lo:hi], w[lo:hi], alpha, nperm, kmax, nmin, min_width, s, lo, hi, stats, sbdry)
        if not cpts:
            ends.append(hi)
        else:
            bounds = [lo] + [lo + c for c in cpts] + [hi]
            for a, b in zip(bounds[:-1], bounds[1:]):
                stack.append((a, b))
    return sorted(ends)


# ---------------------------------------------------------------------------------------------
# CBS.R pre-/post-processing (include/CBS.R:21-132) and exec_cbs (predict_tools.py:242-275)
# ---------------------------------------------------------------------------------------------
def cbs_r(results_r, results_w, ref_gender, alpha, binsize, seed=None, nperm=10000, segmenter=None):
    """Restates CBS.R: returns a list of dicts {"chr" (1-based), "s" (0-based), "e" (exclusive), "r"}.
    ``segmenter(y, w, chrom) -> ends`` defaults to ``segment_chromosome`` above."""
    nchr = 24 if ref_gender == "M" else 23  # CBS.R:30-34
    seed_i = 0 if seed is None else int(seed)
    if segmenter is None:
        def segmenter(yy, ww, c):
            return segment_chromosome(yy, ww, alpha=alpha, nperm=nperm, seed=seed_i, chrom=c)
    na_thresh = int((binsize / 2000000.0) ** -1)  # CBS.R:95
    out = []
    for c in range(nchr):
        ratio = np.asarray(results_r[c], dtype=np.float64)
        wts = np.asarray(results_w[c], dtype=np.float64).copy()
        na = ratio == 0  # CBS.R:41
        wts[wts == 0] = 1.0  # CBS.R:42: 1^-99 == 1 in R
        if na.all():  # CBS.R:56-63
            continue
        keep = np.flatnonzero(~na)
        ends = segmenter(ratio[keep], wts[keep], c)
        starts = [0] + list(ends[:-1])
        for a, b in zip(starts, ends):
            # DNAcopy reports loc.start / loc.end = positions (1-based x) of the first / last point
            start_i, end_i = int(keep[a]) + 1, int(keep[b - 1]) + 1
            seg_na = na[start_i - 1:end_i]
            d = np.diff(seg_na.astype(np.int8))
            start_pos = np.flatnonzero(d == 1) + start_i  # CBS.R:92 (which(...) + start.i - 1, 1-based which)
            end_pos = np.flatnonzero(d == -1) + start_i  # CBS.R:93
            sel = (end_pos - start_pos) > na_thresh
            start_pos, end_pos = start_pos[sel], end_pos[sel]
            inv_start = np.concatenate([[start_i], end_pos])
            inv_end = np.concatenate([start_pos, [end_i]])
            ok = (inv_end - inv_start) > 0
            for s1, e1 in zip(inv_start[ok], inv_end[ok]):
                yy = ratio[s1 - 1:e1]
                ww = wts[s1 - 1:e1]
                m = yy != 0
                r = float(np.sum(yy[m] * ww[m]) / np.sum(ww[m])) if m.any() else float("nan")  # CBS.R:122-127
                out.append({"chr": c + 1, "s": int(s1) - 1, "e": int(e1), "r": r})  # CBS.R:129
    return out


def exec_cbs_segments(results_r, results_w, ref_gender, alpha, binsize, seed=None, nperm=10000):
    """exec_cbs minus the z-scores: [[chr (0-based), s, e, r], ...] (predict_tools.py:266-275)."""
    return [[d["chr"] - 1, d["s"], d["e"], d["r"]] for d in
            cbs_r(results_r, results_w, ref_gender, alpha, binsize, seed, nperm)]
