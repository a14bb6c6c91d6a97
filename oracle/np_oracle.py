"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the WisecondorX numeric hot path.

This is the CPU oracle the CUDA path is checked against (tests/, __graft_entry__.smoke(),
bench.py's cpu_baseline / --impl reference leg).  It is NOT a product path: nothing under
``wisecondorx_b200/`` imports it.

Pinning: every function here is checked against the live reference functions (imported from
/root/reference in the build container) by ``tests/test_oracle_pin.py`` and against the
committed golden vectors under ``tests/golden/`` (generated from the live reference by
``tests/golden/make_golden.py``).  All file:line citations are relative to
/root/reference/src/wisecondorx/.

The restatement is vectorised (argsort instead of the reference's bisect loop, masked array
arithmetic instead of per-bin Python loops); equivalences are noted per function.
"""
from __future__ import annotations

import math

import numpy as np

# --------------------------------------------------------------------------------------
# newref: get_reference (newref_tools.py:155-278)
# --------------------------------------------------------------------------------------


def get_part(partnum: int, outof: int, bincount: int):
    """newref_tools.py:244-247 (partnum is 0-based here as in the reference helper)."""
    return int(bincount / float(outof) * partnum), int(bincount / float(outof) * (partnum + 1))


def chr_of_rows(cum: np.ndarray) -> np.ndarray:
    """Chromosome index of every masked bin given masked_bins_per_chr_cum."""
    n = int(cum[-1])
    return np.searchsorted(np.asarray(cum), np.arange(n), side="right")


def ref_for_bin(x: np.ndarray, row: int, c_start: int, c_end: int, ref_size: int):
    """One target bin of get_ref_for_bins (newref_tools.py:255-278).

    Distance: ``np.sum(np.power(chr_data - row, 2), 1)`` (line 260) -- evaluated with the very
    same NumPy expression so the floating-point order (pairwise summation) is the
    reference's.  Selection: the reference's strict ``<`` against the running max plus
    ``bisect.bisect`` (= bisect_right) keeps, for equal distances, the lower position first
    == first ``ref_size`` entries of a stable argsort.  Values >= 1e10 and NaN
    are never inserted; missing entries stay at -1 / 1e10 (lines 261-262).
    Positions are indexes into the chromosome-excluded array.
    """
    chr_data = np.concatenate((x[:c_start], x[c_end:]))
    d = np.sum(np.power(chr_data - x[row, :], 2), 1)
    order = np.argsort(d, kind="stable")[:ref_size]
    order = order[d[order] < 1e10]  # NaN compares False as in `binVal < cur_max`
    idx = np.full(ref_size, -1, dtype=np.int32)
    dist = np.full(ref_size, 1e10, dtype=np.float64)
    idx[: len(order)] = order
    dist[: len(order)] = d[order]
    return idx, dist


def get_reference_topk(x, per_chr, cum, ref_size, row_begin, row_end):
    """indexes/distances half of get_reference (newref_tools.py:163-209).

    Gonosomal references (len(cum) > 22) give placeholder rows (index 0, distance 1.0) for every
    chromosome other than X/Y (lines 186-191)."""
    x = np.asarray(x)
    per_chr = np.asarray(per_chr, dtype=np.int64)
    cum = np.asarray(cum, dtype=np.int64)
    rows = row_end - row_begin
    idx = np.zeros((rows, ref_size), dtype=np.int32)
    dist = np.ones((rows, ref_size), dtype=np.float64)
    chrom = chr_of_rows(cum)
    gon = len(cum) > 22
    for r in range(row_begin, row_end):
        c = int(chrom[r])
        if gon and c != 22 and c != 23:
            continue
        i, d = ref_for_bin(x, r, int(cum[c] - per_chr[c]), int(cum[c]), ref_size)
        idx[r - row_begin] = i
        dist[r - row_begin] = d
    return idx, dist


def null_ratios(x, idx, row_begin, row_end, sample_ids):
    """Null-ratio loop of get_reference (newref_tools.py:210-224).

    For each chosen sample column (full-length, NOT chromosome-excluded -- quirk SURVEY A.2) and
    each target bin b: log2(col[b] / median(col[idx[b]])); -1 indexes wrap to the last bin.
    ``sample_ids`` is the host-side ``random.sample(range(S), min(S, 100))`` draw (line 215)."""
    x = np.asarray(x)
    out = np.zeros((row_end - row_begin, len(sample_ids)), dtype=np.float64)
    with np.errstate(all="ignore"):
        for m, s in enumerate(sample_ids):
            col = x[:, s]
            ref = col[idx]  # [rows, k] fancy gather, negative wraps
            out[:, m] = np.log2(col[row_begin:row_end] / np.median(ref, axis=1))
    return out


def get_reference(x, per_chr, cum, ref_size, part, split_parts, sample_ids):
    """Full boundary function get_reference (newref_tools.py:155-224); ``part`` is 1-based.
    ``sample_ids`` replaces the internal unseeded random.sample draw."""
    start, end = get_part(part - 1, split_parts, int(cum[-1]))
    idx, dist = get_reference_topk(x, per_chr, cum, ref_size, start, end)
    nr = null_ratios(x, idx, start, end, sample_ids)
    return idx, dist, nr


# --------------------------------------------------------------------------------------
# newref: prep (newref_tools.py:110-147, newref_control.py:38-58)
# --------------------------------------------------------------------------------------


def normalize_and_mask(samples, chrs, mask):
    """newref_tools.py:110-129: stack per-chromosome counts to [bins, S] (zero padded to the
    longest sample), divide each column by its total, keep masked rows."""
    cols = []
    lens = [max(len(s[str(c)]) for s in samples) for c in chrs]
    for s in samples:
        parts = []
        for c, ln in zip(chrs, lens):
            v = np.zeros(ln, dtype=float)
            a = np.asarray(s[str(c)])
            v[: len(a)] = a
            parts.append(v)
        cols.append(np.concatenate(parts))
    all_data = np.stack(cols, axis=1)
    all_data = all_data / np.sum(all_data, 0)
    return all_data[np.asarray(mask, dtype=bool), :]


def pca_fit_exact(t_data, ncomp=5):
    """The PCA model ``sklearn.decomposition.PCA(n_components=5).fit`` approximates
    (newref_tools.py:140-141): mean over samples and the top right singular vectors of the
    centred [S, N] matrix.  Computed through the S x S Gram matrix (exact to ~1e-13 vs full
    SVD).  Sign convention = sklearn's ``svd_flip(u, v, u_based_decision=False)``:
    the largest-|.| entry of each component row is positive."""
    mean = np.mean(t_data, axis=0)
    xc = t_data - mean
    g = xc @ xc.T
    w, u = np.linalg.eigh(g)
    order = np.argsort(w)[::-1][:ncomp]
    w = w[order]
    u = u[:, order]
    comps = (u.T @ xc) / np.sqrt(w)[:, None]
    piv = np.argmax(np.abs(comps), axis=1)
    signs = np.sign(comps[np.arange(ncomp), piv])
    comps = comps * signs[:, None]
    return comps, mean


def train_pca(ref_data, pcacomp=5):
    """newref_tools.py:138-147: corrected = t / inverse_transform(transform(t)), transposed back."""
    t = ref_data.T
    comps, mean = pca_fit_exact(t, pcacomp)
    transformed = (t - mean) @ comps.T
    inversed = transformed @ comps + mean
    return (t / inversed).T, comps, mean


def pca_distance_filter(pca_corrected):
    """newref_control.py:40-46 -> boolean bad-bin mask and the cutoff."""
    med = np.median(pca_corrected, axis=0)
    d = np.sum((pca_corrected - med) ** 2, axis=1)
    mad = np.median(np.abs(d - np.median(d)))
    cutoff = max(np.median(d) + 10 * mad, 5.0)
    return d > cutoff, cutoff, d


# --------------------------------------------------------------------------------------
# predict: normalize (predict_control.py:21-39, predict_tools.py:32-155)
# --------------------------------------------------------------------------------------

Z_MASK = 2.3263478740408408  # scipy.stats.norm.ppf(0.99), predict_tools.py:104


def coverage_normalize_and_mask(sample, bins_per_chr, mask):
    """predict_tools.py:32-48: pad/truncate each chromosome to the reference's bin count,
    divide by the grand total, apply the mask."""
    parts = []
    for c, n in enumerate(bins_per_chr):
        v = np.zeros(int(n), dtype=float)
        a = np.asarray(sample[str(c + 1)])
        m = min(int(n), len(a))
        v[:m] = a[:m]
        parts.append(v)
    all_data = np.concatenate(parts)
    all_data = all_data / np.sum(all_data)
    return all_data[np.asarray(mask, dtype=bool)]


def project_pc(sample_data, comps, mean):
    """predict_tools.py:56-65: x / ((x - mu) C^T C + mu)."""
    t = (sample_data - mean) @ comps.T
    return sample_data / (t @ comps + mean)


def get_weights(distances):
    """predict_tools.py:152-155: 1 / mean(sqrt(row))."""
    return 1.0 / np.mean(np.sqrt(distances), axis=1)


def get_optimal_cutoff(distances, repeats):
    """predict_tools.py:74-82: iterated mean + 3 * population std of the distances below the
    running cutoff."""
    cutoff = float("inf")
    for _ in range(repeats):
        sel = distances[distances < cutoff]
        cutoff = np.average(sel) + 3 * np.std(sel)
    return cutoff


def normalize_once(test_data, test_copy, idx, dist, per_chr, cum, cutoff, ct, cp):
    """predict_tools.py:111-142.  Per target bin i (chromosome >= cp): references are
    chr_excluded(test_copy)[idx[i, dist[i] < cutoff]] with negatives dropped; z uses the
    population std, r the median, n the count."""
    n_out = int(cum[-1]) - ct
    z = np.zeros(n_out)
    r = np.zeros(n_out)
    n = np.zeros(n_out)
    with np.errstate(all="ignore"):
        for c in range(cp, len(per_chr)):
            s, e = int(cum[c] - per_chr[c]), int(cum[c])
            if e <= s:
                continue
            chr_data = np.concatenate((test_copy[:s], test_copy[e:]))
            ii = idx[s:e]
            ref = chr_data[ii]  # negative indexes wrap like the reference's fancy index
            keep = (dist[s:e] < cutoff) & (ref >= 0)
            cnt = keep.sum(axis=1)
            refm = np.where(keep, ref, np.nan)
            mean = np.nansum(refm, axis=1) / cnt
            var = np.nansum((refm - mean[:, None]) ** 2, axis=1) / cnt
            med = np.nanmedian(refm, axis=1)
            med = np.where(cnt > 0, med, np.nan)
            z[s - ct:e - ct] = (test_data[s:e] - mean) / np.sqrt(var)
            r[s - ct:e - ct] = test_data[s:e] / med
            n[s - ct:e - ct] = cnt
    return z, r, n


def normalize_repeat(test_data, idx, dist, per_chr, cum, cutoff, ct, cp):
    """predict_tools.py:94-108: three passes, masking |z| >= norm.ppf(0.99) with -1 between."""
    test_copy = np.copy(test_data)
    z = r = n = None
    for _ in range(3):
        z, r, n = normalize_once(test_data, test_copy, idx, dist, per_chr, cum, cutoff, ct, cp)
        with np.errstate(all="ignore"):
            test_copy[ct:][np.abs(z) >= Z_MASK] = -1
    with np.errstate(all="ignore"):
        m_lr = np.nanmedian(np.log2(r))
        m_z = np.nanmedian(z)
    return z, r, n, m_lr, m_z


def normalize(sample, ref, ref_gender, maskrepeats=5):
    """predict_control.py:21-39.  ``ref`` is a dict-like with the reference .npz keys."""
    if ref_gender == "A":
        ap, cp, ct = "", 0, 0
    else:
        ap, cp = ".{}".format(ref_gender), 22
        ct = int(ref["masked_bins_per_chr_cum" + ap][cp - 1])
    x = coverage_normalize_and_mask(sample, ref["bins_per_chr" + ap], ref["mask" + ap])
    x = project_pc(x, ref["pca_components" + ap], ref["pca_mean" + ap])
    w = get_weights(ref["distances" + ap])[ct:]
    cutoff = get_optimal_cutoff(ref["distances"], maskrepeats)  # always autosomal (A.6)
    z, r, n, m_lr, m_z = normalize_repeat(
        x, ref["indexes" + ap], ref["distances" + ap], ref["masked_bins_per_chr" + ap],
        ref["masked_bins_per_chr_cum" + ap], cutoff, ct, cp)
    return r, z, w, n, m_lr, m_z


# --------------------------------------------------------------------------------------
# predict: between-sample segment z-score (overall_tools.py:88-119)
# --------------------------------------------------------------------------------------


def get_z_score(results_c, results_nr, results_r, results_w):
    """overall_tools.py:88-119.  results_nr[chr] is a per-bin list of null-ratio rows ([M]) or
    scalars 0 for unmasked bins; bins with ratio 0 are dropped; per null column the weighted
    mean over finite entries; z = (segment ratio - mean) / population std, clipped to +-1000;
    the string "nan" when undefined."""
    zs = []
    for seg in results_c:
        c, s, e, ratio = seg[0], seg[1], seg[2], seg[3]
        rr = np.asarray(results_r[c][s:e], dtype=float)
        keep = rr != 0
        rows = [results_nr[c][s + i] for i in range(e - s) if keep[i]]
        w = np.asarray(results_w[c][s:e], dtype=float)[keep]
        if len(rows) == 0:
            zs.append("nan")
            continue
        nr = np.array(rows, dtype=float)  # [bins, M]
        fin = np.isfinite(nr)
        with np.errstate(all="ignore"):
            num = np.sum(np.where(fin, nr, 0.0) * w[:, None], axis=0)
            den = np.sum(np.where(fin, w[:, None], 0.0), axis=0)
            null_seg = num / den
            ok = np.isfinite(null_seg) & (fin.sum(axis=0) > 0)
            vals = null_seg[ok]
            if len(vals) == 0:
                zs.append("nan")
                continue
            null_mean = np.mean(vals)
            null_sd = np.std(vals)
            z = (ratio - null_mean) / null_sd
        z = min(z, 1000)
        z = max(z, -1000)
        if math.isnan(null_mean) or math.isnan(null_sd):
            z = "nan"
        zs.append(z)
    return zs


# --------------------------------------------------------------------------------------
# newref: tool_newref_prep (newref_control.py:24-80) and get_mask (newref_tools.py:77-102)
# --------------------------------------------------------------------------------------


def get_mask(samples):
    """newref_tools.py:77-102: bins whose summed depth-normalised coverage exceeds 5 % of the median
    over the non-empty bins.  Returns (mask over the concatenated 24 chromosomes, bins_per_chr)."""
    lens = [max(len(s[str(c)]) for s in samples) for c in range(1, 25)]
    all_data = normalize_and_mask(samples, range(1, 25), np.ones(int(sum(lens)), dtype=bool))
    sum_per_bin = np.sum(all_data, 1)
    median_cov = np.median(sum_per_bin[sum_per_bin > 0])
    return sum_per_bin > (0.05 * median_cov), lens


def tool_newref_prep(samples, gender, mask, bins_per_chr, pcacomp=5):
    """newref_control.py:24-80 with the exact PCA in place of sklearn's randomized solver.  `mask` is edited IN
    PLACE through a view exactly like the reference (:33, :51-54), so a caller that passes its
    total mask sees the removed bins in the later passes.  Returns the dict the reference saves (:68-80)."""
    last_chr = {"A": 22, "F": 23}.get(gender, 24)
    bins_per_chr = list(bins_per_chr[:last_chr])
    mask = mask[: int(np.sum(bins_per_chr))]
    chrs = range(1, last_chr + 1)
    masked = normalize_and_mask(samples, chrs, mask)
    corrected, comps, mean = train_pca(masked, pcacomp)
    bad, cutoff, _ = pca_distance_filter(corrected)
    if np.any(bad):
        mask[np.where(mask)[0][bad]] = False
        masked = normalize_and_mask(samples, chrs, mask)
        corrected, comps, mean = train_pca(masked, pcacomp)
    offs = np.concatenate([[0], np.cumsum(bins_per_chr)]).astype(int)
    mbpc = [int(np.sum(mask[offs[i]:offs[i + 1]])) for i in range(len(bins_per_chr))]
    return {"mask": mask.copy(), "bins_per_chr": np.array(bins_per_chr), "masked_bins_per_chr": np.array(mbpc),
            "masked_bins_per_chr_cum": np.cumsum(mbpc), "pca_components": comps, "pca_mean": mean,
            "pca_corrected_data": corrected, "n_removed": int(np.sum(bad)), "cutoff": float(cutoff)}


# --------------------------------------------------------------------------------------
# predict: result assembly and post-processing (main.py:242-271, predict_control.py:49-63,
# predict_tools.py:163-233)
# --------------------------------------------------------------------------------------


def assemble_results(aut, gon, nr_aut, nr_gon, minrefbins, mask, bins_per_chr):
    """main.py:242-271 + get_post_processed_result (predict_control.py:49-63).  aut / gon are the tuples
    (results_r, results_z, results_w, ref_sizes, m_lr, m_z) of the two `normalize` calls.  Returns per-chromosome
    lists of float arrays for r / z / w (0 = no data) and the ref_sizes vector."""
    r = np.append(aut[0], gon[0])
    z = np.append(aut[1], gon[1]) - aut[5]
    with np.errstate(all="ignore"):
        w = np.append(aut[2] * np.nanmean(gon[2]), gon[2] * np.nanmean(aut[2]))
        w = w / np.nanmean(w)
    if np.isnan(w).any() or np.isinf(w).any():
        w = np.ones(len(w))
    ref_sizes = np.append(aut[3], gon[3])
    mask = np.asarray(mask, dtype=bool)
    offs = np.concatenate([[0], np.cumsum(bins_per_chr)]).astype(int)
    out = {}
    for key, val in (("results_r", r), ("results_z", z), ("results_w", w)):
        val = np.array(val, dtype=float)
        val[ref_sizes < minrefbins] = 0  # predict_control.py:50-51
        full = np.zeros(len(mask))
        full[mask] = val[: int(mask.sum())]  # inflate_results, predict_tools.py:163-170
        out[key] = [full[offs[c]:offs[c + 1]] for c in range(len(bins_per_chr))]
    return out, ref_sizes


def log_trans(results, log_r_median):
    """predict_tools.py:180-193."""
    for c in range(len(results["results_r"])):
        with np.errstate(all="ignore"):
            r = np.log2(results["results_r"][c])
        bad = ~np.isfinite(r)
        r[bad] = 0
        results["results_z"][c][bad] = 0
        results["results_w"][c][bad] = 0
        r[r != 0] -= log_r_median
        results["results_r"][c] = r


def per_bin_stats(indexes, distances):
    """ref_qc.py:22-38: mean and max distance and number of reference bins of every target bin."""
    d = np.asarray(distances, dtype=float)
    return np.mean(d, axis=1), np.max(d, axis=1), np.full(len(d), np.asarray(indexes).shape[1], dtype=int)


def apply_blacklist(results, bed_text, binsize):
    """predict_tools.py:202-233: bins [int(s / binsize), int(e / binsize) + 1) of every BED line are blanked in
    r / z / w; chrY lines are skipped when the results hold 23 chromosomes; out-of-range positions are ignored."""
    for line in bed_text.splitlines():
        if not line.strip():
            continue
        name, s, e = line.strip().split("\t")
        name = name[3:] if name[:3] == "chr" else name
        c = int({"X": "23", "Y": "24"}.get(name, name)) - 1
        if len(results["results_r"]) < 24 and c == 23:
            continue
        lo = max(0, int(int(s) / binsize))
        hi = min(len(results["results_r"][c]), int(int(e) / binsize) + 1)
        for key in ("results_r", "results_z", "results_w"):
            results[key][c][lo:hi] = 0
