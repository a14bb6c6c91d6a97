"""Reference QC run at the end of `newref` (mirror of the reference's ref_qc.py; main.py:134-135 calls
`qc_reference(args.outfile)` -- without importing it, a NameError in v1.2.10 that this package does not reproduce).
Same metrics, thresholds, log lines and return code (0 PASS / 1 WARN / 2 FAIL); the per-bin statistics
(ref_qc.py:22-38, a Python loop over every bin) are three vectorised reductions over the [bins, refsize] arrays."""
from __future__ import annotations

import logging
from pathlib import Path

import numpy as np

MINREFBINS = 150
OUTLIER_N_SIGMA = 3


def _get_gender_suffixes(keys):
    out = [sfx for sfx in (".F", ".M") if "bins_per_chr" + sfx in keys]
    if "bins_per_chr" in keys and not out:
        out.append("")
    return out


def _row_blocks(fn, d):
    """fn(block, axis=1) over row blocks of d on a thread pool (NumPy releases the GIL; per-row results do not depend on
    the blocking)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    n = len(d)
    nthreads = max(1, min(16, len(os.sched_getaffinity(0))))
    if n < 1 << 15 or nthreads == 1:
        return fn(d, axis=1)
    step = max(4096, -(-n // (4 * nthreads)))
    out = np.empty(n, dtype=d.dtype)
    with ThreadPoolExecutor(nthreads) as pool:
        list(pool.map(lambda a: fn(d[a:a + step], axis=1, out=out[a:a + step]), range(0, n, step)))
    return out


def compute_per_bin_stats(indexes, distances, need_max=True):
    """mean / max distance and number of reference bins per target bin (ref_qc.py:22-38).  need_max=False skips the
    maximum (compute_metrics does not use it; 0.5 GB per gonosomal set at 15 kb)."""
    d = np.asarray(distances, dtype=float)
    idx = np.asarray(indexes)
    if d.ndim == 1:
        d, idx = d[:, None], idx[:, None]
    n = len(idx)
    if d.shape[1] == 0:
        return np.full(n, np.nan), np.full(n, np.nan), np.zeros(n, dtype=int)
    return _row_blocks(np.mean, d), (_row_blocks(np.max, d) if need_max else None), np.full(n, idx.shape[1], dtype=int)


def _chrY_metrics(ref, suf, mean_d, n_refs, cutoff_outlier):
    if suf != ".M" or "masked_bins_per_chr_cum" + suf not in ref:
        return None
    cum = np.atleast_1d(ref["masked_bins_per_chr_cum" + suf])
    if len(cum) < 24:
        return None
    start, end = int(cum[22]), int(cum[23])
    if start >= end:
        return {"n_bins": 0}
    m, r = mean_d[start:end], n_refs[start:end]
    valid = np.isfinite(m)
    if not valid.any():
        return {"n_bins": end - start, "n_valid": 0, "mean_of_means": np.nan}
    return {"n_bins": end - start, "n_valid": int(valid.sum()), "mean_of_means": float(np.mean(m[valid])),
            "std_of_means": float(np.std(m[valid])), "n_mean_outlier": int(np.sum(m[valid] >= cutoff_outlier)),
            "n_low_refs": int(np.sum(r < MINREFBINS))}


def compute_metrics(ref, suf):
    """ref_qc.py:69-104."""
    if "indexes" + suf not in ref or "distances" + suf not in ref:
        return None
    indexes, distances = ref["indexes" + suf], ref["distances" + suf]
    n_bins = len(indexes)
    if n_bins == 0:
        return {"n_bins": 0}
    mean_d, _, n_refs = compute_per_bin_stats(indexes, distances, need_max=False)
    valid = np.isfinite(mean_d)
    n_valid = int(valid.sum())
    if n_valid == 0:
        return {"n_bins": n_bins, "n_valid": 0}
    mean_of_means = float(np.mean(mean_d[valid]))
    std_of_means = float(np.std(mean_d[valid]))
    cutoff = mean_of_means + OUTLIER_N_SIGMA * std_of_means
    n_out = int(np.sum(mean_d[valid] >= cutoff))
    return {"n_bins": n_bins, "n_valid": n_valid, "mean_of_means": mean_of_means, "std_of_means": std_of_means,
            "n_mean_outlier": n_out, "outlier_pct": 100.0 * n_out / n_valid, "n_low_refs": int(np.sum(n_refs < MINREFBINS)),
            "chrY": _chrY_metrics(ref, suf, mean_d, n_refs, cutoff)}


def _verdict(m, male):
    """ref_qc.py:107-137 (_verdict_f / _verdict_m)."""
    if m is None or m.get("n_valid", 0) == 0:
        return "FAIL", "no data"
    if m["n_low_refs"] > 0:
        return "WARN", f"n_refs<{MINREFBINS} in {m['n_low_refs']} bins"
    if not male:
        if m["std_of_means"] > 10:
            return "FAIL", f"std(per-bin mean dist) = {m['std_of_means']:.2f} (high)"
        if m["std_of_means"] > 2:
            return "WARN", f"std(per-bin mean dist) = {m['std_of_means']:.2f}"
    else:
        if m["mean_of_means"] > 10:
            return "FAIL", f"mean(per-bin mean dist) = {m['mean_of_means']:.2f} (heavy tail)"
        if m["mean_of_means"] > 2:
            return "WARN", f"mean(per-bin mean dist) = {m['mean_of_means']:.2f}"
        cy = m.get("chrY")
        if cy and cy.get("n_valid", 0) > 0 and np.isfinite(cy.get("mean_of_means", np.nan)):
            if cy["mean_of_means"] > 100:
                return "FAIL", f"chrY mean distance = {cy['mean_of_means']:.1f} (very poor chrY)"
            if cy["mean_of_means"] > 5:
                return "WARN", f"chrY mean distance = {cy['mean_of_means']:.1f}"
    if m["outlier_pct"] > 1:
        return "WARN", f"outlier bins = {m['outlier_pct']:.2f}%"
    return "PASS", ""


def qc_reference(npz_path, ref=None):
    """Drop-in for ref_qc.qc_reference (ref_qc.py:140-222).  `ref` (optional): the arrays of the reference that was
    just written, so `newref` does not re-inflate its own output."""
    if ref is None:
        npz = Path(npz_path).resolve()
        if not npz.exists():
            logging.error(f"QC check skipped: file not found: {npz}")
            return 2
        from . import npz_io
        ref = npz_io.load_npz(str(npz))
    keys = list(ref.keys())
    try:
        binsize = int(np.atleast_1d(ref["binsize"])[0])
    except Exception:
        binsize = None
    suffixes = _get_gender_suffixes(keys)
    if not suffixes:
        logging.error("QC failed: no bins_per_chr / bins_per_chr.F / bins_per_chr.M in npz")
        return 2
    logging.info("Starting ref-QC for file: {}".format(Path(npz_path).resolve()))
    logging.info("Reference binsize: {} bp".format(binsize) if binsize else "Reference binsize: (unknown)")
    worst = 0
    for suf in suffixes:
        label = {".F": "F", ".M": "M"}.get(suf, "A")
        m = compute_metrics(ref, suf)
        if m is None:
            logging.warning(f"[{label}] no indexes/distances — skip")
            continue
        if m.get("n_valid", 0) == 0:
            logging.error(f"[{label}] n_bins={m['n_bins']}, n_valid=0 — FAIL")
            worst = max(worst, 2)
            continue
        verdict, msg = _verdict(m, label == "M")
        worst = max(worst, {"FAIL": 2, "WARN": 1}.get(verdict, 0))
        log = {"FAIL": logging.error, "WARN": logging.warning}.get(verdict, logging.info)
        log(f"[{label}] n_bins={m['n_bins']}, mean(dist)={m['mean_of_means']:.4f}, std(dist)={m['std_of_means']:.4f}, "
            f"outliers={m['n_mean_outlier']} ({m['outlier_pct']:.2f}%), n_refs<{MINREFBINS}={m['n_low_refs']}")
        cy = m.get("chrY")
        if cy and cy.get("n_valid", 0) > 0:
            log(f"       chrY: n_bins={cy['n_bins']}, mean={cy['mean_of_means']:.4f}, std={cy['std_of_means']:.4f}, "
                f"outliers={cy['n_mean_outlier']}, n_refs<{MINREFBINS}={cy['n_low_refs']}")
        log(f"         -> {verdict}" + (f": {msg}" if msg else ""))
    if worst == 0:
        logging.info("QC Overall Verdict: PASS")
    elif worst == 1:
        logging.warning("QC Overall Verdict: WARN (review metrics above)")
    else:
        logging.error("QC Overall Verdict: FAIL (ref may cause poor predictions; consider rebuilding or more samples)")
    return worst
