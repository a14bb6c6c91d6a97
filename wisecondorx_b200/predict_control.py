"""Host-side mirror of the reference's predict controller (predict_control.py)."""
from __future__ import annotations

import numpy as np

from . import predict_tools


def normalize(args, sample, ref_file, ref_gender, engine: predict_tools.PredictEngine | None = None):
    """Drop-in for predict_control.normalize (reference predict_control.py:21-39): coverage
    normalisation, between-sample (PCA) normalisation and the three within-sample passes.
    Returns (results_r, results_z, results_w, ref_sizes, m_lr, m_z)."""
    r, z, w, n, m_lr, m_z = normalize_batch(args, [sample], ref_file, ref_gender, engine)
    return r[0], z[0], w, n[0], float(m_lr[0]), float(m_z[0])


def normalize_batch(args, samples, ref_file, ref_gender, engine: predict_tools.PredictEngine | None = None):
    """`normalize` for a batch of samples against one reference (BASELINE config 5): the reference
    arrays are read once per batch.  Returns arrays with a leading batch axis (weights are shared)."""
    eng = engine or predict_tools.default_engine()
    if ref_gender == "A":
        ap, cp, ct = "", 0, 0
    else:
        ap = ".{}".format(ref_gender)
        cp = 22
        ct = int(ref_file["masked_bins_per_chr_cum{}".format(ap)][cp - 1])
    results_w = eng.get_weights(ref_file, ap)[ct:]
    optimal_cutoff = eng.get_optimal_cutoff(ref_file, args.maskrepeats)
    z, r, n, m_lr, m_z = eng.normalize_set(samples, ref_file, ap, optimal_cutoff, cp, ct)
    return r, z, results_w, n, m_lr, m_z


def inflate_results(results, mask):
    """predict_tools.inflate_results (reference :163-170), vectorised."""
    mask = np.asarray(mask, dtype=bool)
    out = np.zeros(len(mask), dtype=np.asarray(results).dtype)
    out[mask] = results
    return out


def segment_batch(r, w, nref, m_lr, chr_offsets, minrefbins=150, alpha=1e-4, nperm=10000, seed=0, ctx=None):
    """CBS of a normalised batch in one call: the log2 ratios of every (sample, chromosome) as one series each
    (bins without enough reference bins or without signal are dropped like NA bins in CBS.R:41).  Returns the list
    of segment-end arrays in (sample, chromosome) order."""
    from . import cbs
    offs = np.asarray(chr_offsets, dtype=np.int64)
    nchr = len(offs) - 1
    series = []
    for i in range(r.shape[0]):
        with np.errstate(all="ignore"):
            lr = np.log2(r[i]) - m_lr[i]
        ok = np.isfinite(lr) & (nref[i] >= minrefbins) & (lr != 0)
        for c in range(nchr):
            m = ok[offs[c]:offs[c + 1]]
            series.append((lr[offs[c]:offs[c + 1]][m], w[offs[c]:offs[c + 1]][m]))
    return cbs.segment_series(series, [i % nchr for i in range(len(series))], alpha=alpha, nperm=nperm, seed=seed, ctx=ctx)
