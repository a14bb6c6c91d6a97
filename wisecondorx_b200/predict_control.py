"""Host-side mirror of the reference's predict controller (predict_control.py)."""
from __future__ import annotations

import logging
import warnings

import numpy as np

from . import _lib, predict_tools
from .predict_tools import _ptr, map_threads as _map_threads
from .overall_tools import gender_correct, scale_sample


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
    w = np.asarray(w, dtype=np.float64)

    def one(i):  # the kept bins of the whole sample at once; the chromosomes are ranges of them
        with np.errstate(all="ignore"):
            lr = np.log2(r[i]) - m_lr[i]
        ok = np.isfinite(lr)
        ok &= nref[i] >= minrefbins
        ok &= lr != 0
        cols = np.flatnonzero(ok)
        cols = cols[(cols >= offs[0]) & (cols < offs[-1])]
        return lr[cols], w[cols], np.diff(np.searchsorted(cols, offs))

    parts = _map_threads(one, range(r.shape[0]))
    if not parts:
        return []
    off = np.concatenate([[0], np.cumsum(np.concatenate([p[2] for p in parts]))]).astype(np.int64)
    ends, nseg = cbs._segment_flat(np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]), off,
                                   np.tile(np.arange(nchr, dtype=np.int32), len(parts)), alpha, nperm, seed, ctx)
    cut = np.concatenate([[0], np.cumsum(nseg)]).astype(np.int64)
    return [ends[cut[s]:cut[s + 1]].copy() for s in range(len(off) - 1)]


# ---------------------------------------------------------------------------------------------
# the whole numeric flow of `predict` (reference main.py:168-290 up to the output writers) for a batch of samples
# ---------------------------------------------------------------------------------------------
def _y_fraction(sample):
    return float(np.sum(sample["24"])) / float(np.sum([np.sum(sample[x]) for x in sample.keys()]))


def predict_gender(sample, trained_cutoff):
    return "M" if _y_fraction(sample) > trained_cutoff else "F"  # reference predict_tools.py:17-24


def get_post_processed_result(minrefbins, result, ref_sizes, mask, bins_per_chr):
    """Zeroes bins with fewer than minrefbins reference bins, unmasks and splits per chromosome
    (reference predict_control.py:49-63 + predict_tools.py:163-170), vectorised."""
    result = np.array(result, dtype=float)
    result[ref_sizes < minrefbins] = 0
    full = np.zeros(len(mask), dtype=float)
    mask_b = np.asarray(mask, dtype=bool)
    cnt = int(np.sum(mask_b))
    # the reference's inflate loop hands out results[j] to the j-th kept bin (predict_tools.py:163-170): surplus
    # results are ignored, missing ones raise.  The counts differ when the gonosomal pass of newref removed
    # autosomal bins after the autosomal snapshot -- preserved, not fixed.
    if len(result) < cnt:
        raise IndexError("list index out of range")
    full[mask_b] = result[:cnt]
    offs = np.concatenate([[0], np.cumsum(bins_per_chr)]).astype(int)
    return [full[offs[c]:offs[c + 1]] for c in range(len(bins_per_chr))]


def log_trans(results, log_r_median):
    """log2 of the ratios; non-finite entries blank r, z and w; the median log-ratio is subtracted
    from every non-zero entry (reference predict_tools.py:180-193)."""
    for c in range(len(results["results_r"])):
        with np.errstate(all="ignore"):
            r = np.log2(results["results_r"][c])
        bad = ~np.isfinite(r)
        r[bad] = 0
        results["results_z"][c][bad] = 0
        results["results_w"][c][bad] = 0
        nz = r != 0
        r[nz] = r[nz] - log_r_median
        results["results_r"][c] = r


def apply_blacklist(path, binsize, results):
    """Blanks the bins overlapping the BED intervals of --blacklist (reference predict_tools.py:202-233)."""
    for line in open(path):
        chr_name, s, e = line.strip().split("\t")[:3]
        chr_name = chr_name[3:] if chr_name[:3] == "chr" else chr_name
        c = {"X": 23, "Y": 24}.get(chr_name, None) or int(chr_name)
        c -= 1
        if len(results["results_r"]) < 24 and c == 23:
            continue
        lo, hi = max(0, int(int(s) / binsize)), min(len(results["results_r"][c]), int(int(e) / binsize) + 1)
        for key in ("results_r", "results_z", "results_w"):
            results[key][c][lo:hi] = 0


def resolve_genders(args, sample, ref_file):
    """(sample after gender_correct, gender, ref_gender) as reference main.py:188-230."""
    gender = predict_gender(sample, ref_file["trained_cutoff"])
    if not ref_file["is_nipt"]:
        if args.gender:
            gender = args.gender
        sample = gender_correct(sample, gender)
        ref_gender = gender
        if not ref_file["has_male"] and gender == "M":
            logging.warning("This sample is male, whilst the reference is created with fewer than 5 males. "
                            "The female gonosomal reference will be used for X predictions.")
            ref_gender = "F"
        elif not ref_file["has_female"] and gender == "F":
            logging.warning("This sample is female, whilst the reference is created with fewer than 5 females. "
                            "The male gonosomal reference will be used for XY predictions.")
            ref_gender = "M"
    else:
        if args.gender:
            gender = args.gender
        ref_gender = "F"
    return sample, gender, ref_gender


def shared_weights(w_aut, w_gon):
    """The bin weights of the stacked autosomal + gonosomal result (reference main.py:246-247) and whether they are
    numeric (:254-259).  They depend on the reference only: once per (reference, gender), not once per sample."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        w = np.append(w_aut * np.nanmean(w_gon), w_gon * np.nanmean(w_aut))
        w = w / np.nanmean(w)
    ok = not (np.isnan(w).any() or np.isinf(w).any())
    return (w if ok else np.ones(len(w))), ok


def _log2_rows(src, rows, inplace=False):
    """log2 of the rows `rows` of src (NumPy's log2, one row per task: predict_tools.py:182) as a [len(rows), n] matrix;
    inplace: the rows are overwritten and the matrix is src itself when `rows` are all its rows in order (no new pages)."""
    same = inplace and len(rows) == len(src) and np.array_equal(rows, np.arange(len(src)))
    out = src if same else np.empty((len(rows), src.shape[1]))

    def one(j):
        with np.errstate(all="ignore"):
            np.log2(src[rows[j]], out=out[j])

    _map_threads(one, range(len(rows)))
    return out


def assemble_batch(args, aut, aut_rows, gon, nr, ref_file, ref_gender, genders, n_reads, weights=None, consume=False):
    """Result assembly (reference main.py:232-271) of the samples of a batch that share the reference gender.
    aut = (r, z [*, n_aut], w [n_aut], ref_sizes [*, n_aut], m_lr, m_z [*]) of the autosomal normalize_batch call and
    aut_rows the rows of it these samples occupy; gon = (r2, z2 [b, n_gon], w2 [n_gon], n2 [b, n_gon]) of their
    gonosomal call; nr = the stacked null ratios of ref_gender; genders / n_reads per sample.  Returns
    [(rem_input, results), ...].  consume: the ratio matrices may be overwritten with their logarithms (the caller
    drops them anyway).

    get_post_processed_result (predict_control.py:49-63), inflate_results (predict_tools.py:163-170) and log_trans
    (predict_tools.py:180-193) run as ONE pass per sample over the bin axis on host threads (wcx_predict_assemble,
    csrc/host_predict.cu) instead of ~100 NumPy calls per sample; the per-chromosome vectors of a sample are views of
    one row, the rows of the batch rows of one matrix (predict_tools.flatten hands them to the CBS and z-score calls
    without a copy)."""
    import os
    sfx = ".{}".format(ref_gender)
    r, z, w, ref_sizes, m_lr, m_z = aut
    r2, z2, w2, n2 = gon[:4]
    b = len(aut_rows)
    w_shared, numeric = weights if weights is not None else shared_weights(w, w2)
    mask = np.ascontiguousarray(ref_file["mask" + sfx], dtype=bool)
    bpc = ref_file["bins_per_chr" + sfx]
    rows = np.ascontiguousarray(aut_rows, dtype=np.int32)
    r, z, ref_sizes, r2, z2, n2 = (np.ascontiguousarray(x, dtype=np.float64).reshape(len(x), -1) for x in (r, z, ref_sizes, r2, z2, n2))
    if len(w_shared) != r.shape[1] + r2.shape[1]:  # the reference's boolean index fails the same way
        raise IndexError("boolean index did not match indexed array")
    lr = _log2_rows(r, rows, consume)
    lr2 = _log2_rows(r2, np.arange(b), consume)
    out_r, out_z, out_w = np.empty((b, len(mask))), np.empty((b, len(mask))), np.empty((b, len(mask)))
    out_i = np.empty((b, len(mask)), dtype=np.int32)
    ml = np.ascontiguousarray(np.asarray(m_lr, dtype=np.float64).reshape(-1)[rows])
    mz = np.ascontiguousarray(np.asarray(m_z, dtype=np.float64).reshape(-1)[rows])
    w_shared = np.ascontiguousarray(w_shared, dtype=np.float64)
    mask_u8 = mask.view(np.uint8)
    rc = _lib.load().wcx_predict_assemble(_ptr(lr), _ptr(z), _ptr(ref_sizes), r.shape[1], _ptr(rows), _ptr(lr2), _ptr(z2), _ptr(n2),
                                          r2.shape[1], _ptr(w_shared), _ptr(ml), _ptr(mz), b, float(args.minrefbins), _ptr(mask_u8),
                                          len(mask), _ptr(out_r), _ptr(out_z), _ptr(out_w), _ptr(out_i),
                                          min(16, max(1, len(os.sched_getaffinity(0)) - 1)))
    if rc == 2:
        # the reference's inflate loop hands out results[j] to the j-th kept bin (predict_tools.py:163-170): surplus
        # results are ignored, missing ones raise.  The counts differ when the gonosomal pass of newref removed
        # autosomal bins after the autosomal snapshot -- preserved, not fixed.
        raise IndexError("list index out of range")
    _lib.check(rc)
    offs = np.concatenate([[0], np.cumsum(bpc)]).astype(int)
    out = []
    for j in range(b):
        if not numeric:
            logging.warning("Non-numeric values found in weights -- reference too small. Circular binary segmentation and "
                            "z-scoring will be unweighted")
        rem_input = {
            "args": args, "binsize": int(ref_file["binsize"]), "n_reads": n_reads[j], "ref_gender": ref_gender, "gender": genders[j],
            "mask": ref_file["mask" + sfx], "bins_per_chr": bpc,
            "masked_bins_per_chr": ref_file["masked_bins_per_chr" + sfx],
            "masked_bins_per_chr_cum": ref_file["masked_bins_per_chr_cum" + sfx],
        }
        results = {key: [val[j, offs[c]:offs[c + 1]] for c in range(len(bpc))]
                   for key, val in (("results_r", out_r), ("results_z", out_z), ("results_w", out_w))}
        results["results_nr"] = {"dense": nr, "inflate": out_i[j]}
        if getattr(args, "blacklist", None):
            apply_blacklist(args.blacklist, rem_input["binsize"], results)
        out.append((rem_input, results))
    return out


def assemble(args, aut, gon, nr, ref_file, ref_gender, gender, n_reads, weights=None):
    """Result assembly of one sample (reference main.py:232-271): aut / gon = (r, z, w, ref_sizes, m_lr, m_z) of the
    two `normalize` calls, nr = the stacked autosomal + gonosomal null ratios of ref_gender."""
    r, z, w, n, m_lr, m_z = aut
    one = lambda x: np.asarray(x, dtype=np.float64).reshape(1, -1)  # noqa: E731
    return assemble_batch(args, (one(r), one(z), w, one(n), [m_lr], [m_z]), [0], (one(gon[0]), one(gon[1]), gon[2], one(gon[3])),
                          nr, ref_file, ref_gender, [gender], [n_reads], weights)[0]


def stacked_null_ratios(ref_file, ref_gender):
    """Autosomal null ratios followed by the gonosomal rows of ref_gender (reference main.py:216-219, :252), one
    dense [rows, M] array (NaN padded when the two sets were built with different numbers of null samples)."""
    nr_aut = ref_file["null_ratios"]
    nr_gon = ref_file["null_ratios.{}".format(ref_gender)][len(nr_aut):]
    if len(nr_gon) and nr_gon.shape[1] == nr_aut.shape[1]:
        return np.concatenate([nr_aut, nr_gon]).astype(np.float64, copy=False)
    m = max(nr_aut.shape[1], nr_gon.shape[1] if len(nr_gon) else 0)
    nr = np.full((len(nr_aut) + len(nr_gon), m), np.nan)
    nr[:len(nr_aut), :nr_aut.shape[1]] = nr_aut
    if len(nr_gon):
        nr[len(nr_aut):, :nr_gon.shape[1]] = nr_gon
    return nr


def predict_batch(args, samples, binsizes, ref_file, engine: predict_tools.PredictEngine | None = None, timings=None):
    """Everything `predict` computes before the output writers, for a batch of raw samples against one reference:
    re-binning, gender, both `normalize` calls (the autosomal one for the whole batch at once, the gonosomal one per
    reference gender), the result assembly, log transform, blacklist, CBS (one device call for every chromosome of
    every sample) and the segment z-scores.  Returns [(rem_input, results), ...] in sample order; results carry
    results_c like the reference's dict after exec_cbs (main.py:283)."""
    import time
    from . import cbs
    eng = engine or predict_tools.default_engine()
    t0 = time.perf_counter()
    def prelude(t):
        sample, bs = t
        reads = int(sum(int(np.sum(sample[x], dtype=np.int64)) for x in sample.keys()))
        sample = scale_sample(dict(sample), int(bs), int(ref_file["binsize"]))
        return resolve_genders(args, sample, ref_file) + (reads,)

    pre = [prelude(t) for t in zip(samples, binsizes)]  # ~25 tiny NumPy calls per sample: a thread pool only adds lock traffic (measured)
    prepared, genders, ref_genders, n_reads = ([p[i] for p in pre] for i in range(4))
    logging.info("Normalizing autosomes ...")
    r, z, w, n, m_lr, m_z = normalize_batch(args, prepared, ref_file, "A", eng)
    logging.info("Normalizing gonosomes ...")
    out = [None] * len(prepared)
    for rg in sorted(set(ref_genders)):
        ids = [i for i, x in enumerate(ref_genders) if x == rg]
        r2, z2, w2, n2, _, _ = normalize_batch(args, [prepared[i] for i in ids], ref_file, rg, eng)
        group = assemble_batch(args, (r, z, w, n, m_lr, m_z), ids, (r2, z2, w2, n2), eng.stacked_null_ratios(ref_file, rg),
                               ref_file, rg, [genders[i] for i in ids], [n_reads[i] for i in ids], consume=True)
        for i, res in zip(ids, group):
            out[i] = res
    if timings is not None:
        timings["normalize_and_assemble"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    logging.info("Executing circular binary segmentation ...")
    if len(out) == 1:
        out[0][1]["results_c"] = cbs.exec_cbs(out[0][0], out[0][1], eng)
    else:
        for (rem, res), rc in zip(out, cbs.exec_cbs_batch([o[0] for o in out], [o[1] for o in out], eng)):
            res["results_c"] = rc
    if timings is not None:
        timings["cbs_and_segment_z"] = time.perf_counter() - t0
    return out
