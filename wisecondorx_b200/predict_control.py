"""Host-side mirror of the reference's predict controller (predict_control.py)."""
from __future__ import annotations

import logging
import warnings

import numpy as np

from . import predict_tools
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
    series = []
    for i in range(r.shape[0]):
        with np.errstate(all="ignore"):
            lr = np.log2(r[i]) - m_lr[i]
        ok = np.isfinite(lr) & (nref[i] >= minrefbins) & (lr != 0)
        for c in range(nchr):
            m = ok[offs[c]:offs[c + 1]]
            series.append((lr[offs[c]:offs[c + 1]][m], w[offs[c]:offs[c + 1]][m]))
    return cbs.segment_series(series, [i % nchr for i in range(len(series))], alpha=alpha, nperm=nperm, seed=seed, ctx=ctx)


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


class _Rows:
    """Output rows of one reference gender: r / z / w [samples, bins] and the bin -> null-ratio-row map, so that the
    per-chromosome vectors of the samples lie back to back (predict_tools.flatten makes one vector of them for the
    CBS and z-score calls without a copy)."""

    def __init__(self, b, mask):
        self.kept = np.flatnonzero(np.asarray(mask, dtype=bool))
        bins = len(mask)
        self.r, self.z, self.w = np.zeros((b, bins)), np.zeros((b, bins)), np.zeros((b, bins))
        self.inflate = np.full((b, bins), -1, dtype=np.int32)


def assemble(args, aut, gon, nr, ref_file, ref_gender, gender, n_reads, weights=None, rows=None, row=0):
    """Result assembly of one sample (reference main.py:232-271): aut / gon = (r, z, w, ref_sizes, m_lr, m_z) of the
    two `normalize` calls, nr = the stacked autosomal + gonosomal null ratios of ref_gender (shared by the batch).

    get_post_processed_result (predict_control.py:49-63), inflate_results (predict_tools.py:163-170) and log_trans
    (predict_tools.py:180-193) run here on the kept bins of the whole genome at once instead of per key and per
    chromosome -- the same elementwise arithmetic, a dozen passes over the bins instead of a hundred NumPy calls per
    sample.  weights: shared_weights() of the reference gender when the caller has it; rows / row: where a batch
    wants the sample's vectors (a _Rows of the gender's mask length)."""
    sfx = ".{}".format(ref_gender)
    rem_input = {
        "args": args, "binsize": int(ref_file["binsize"]), "n_reads": n_reads, "ref_gender": ref_gender, "gender": gender,
        "mask": ref_file["mask" + sfx], "bins_per_chr": ref_file["bins_per_chr" + sfx],
        "masked_bins_per_chr": ref_file["masked_bins_per_chr" + sfx],
        "masked_bins_per_chr_cum": ref_file["masked_bins_per_chr_cum" + sfx],
    }
    results_r, results_z, results_w, ref_sizes, m_lr, m_z = aut
    r2, z2, w2, n2 = gon[:4]
    w_shared, numeric = weights if weights is not None else shared_weights(results_w, w2)
    if not numeric:
        logging.warning("Non-numeric values found in weights -- reference too small. Circular binary segmentation and "
                        "z-scoring will be unweighted")
    mask, bpc = rem_input["mask"], rem_input["bins_per_chr"]
    if rows is None:
        rows = _Rows(1, mask)
    kept = rows.kept
    cnt = len(kept)
    # the reference's inflate loop hands out results[j] to the j-th kept bin (predict_tools.py:163-170): surplus
    # results are ignored, missing ones raise.  The counts differ when the gonosomal pass of newref removed
    # autosomal bins after the autosomal snapshot -- preserved, not fixed.
    if len(results_r) + len(r2) < cnt:
        raise IndexError("list index out of range")
    if len(results_w) + len(w2) != len(results_r) + len(r2):  # the reference's boolean index fails the same way
        raise IndexError("boolean index did not match indexed array")
    low = (np.append(ref_sizes, n2) < args.minrefbins)[:cnt]  # predict_control.py:50-51
    with np.errstate(all="ignore"):
        lr = np.log2(np.append(results_r, r2)[:cnt])  # predict_tools.py:182
    bad = ~np.isfinite(lr)
    bad |= low  # a blanked ratio is 0, its logarithm -inf
    lr[bad] = 0
    np.subtract(lr, m_lr, out=lr, where=lr != 0)  # predict_tools.py:189-191
    z = np.append(results_z, z2)[:cnt] - m_z  # main.py:244
    z[bad] = 0
    w = np.array(w_shared[:cnt], dtype=float)
    w[bad] = 0
    pos = np.arange(cnt, dtype=np.int32)
    pos[low] = -1
    rows.r[row, kept], rows.z[row, kept], rows.w[row, kept], rows.inflate[row, kept] = lr, z, w, pos
    offs = np.concatenate([[0], np.cumsum(bpc)]).astype(int)
    results = {key: [val[row, offs[c]:offs[c + 1]] for c in range(len(bpc))]
               for key, val in (("results_r", rows.r), ("results_z", rows.z), ("results_w", rows.w))}
    results["results_nr"] = {"dense": nr, "inflate": rows.inflate[row]}
    if getattr(args, "blacklist", None):
        apply_blacklist(args.blacklist, rem_input["binsize"], results)
    return rem_input, results


def stacked_null_ratios(ref_file, ref_gender):
    """Autosomal null ratios followed by the gonosomal rows of ref_gender (reference main.py:216-219, :252), one
    dense [rows, M] array (NaN padded when the two sets were built with different numbers of null samples)."""
    nr_aut = ref_file["null_ratios"]
    nr_gon = ref_file["null_ratios.{}".format(ref_gender)][len(nr_aut):]
    m = max(nr_aut.shape[1], nr_gon.shape[1] if len(nr_gon) else 0)
    nr = np.full((len(nr_aut) + len(nr_gon), m), np.nan)
    nr[:len(nr_aut), :nr_aut.shape[1]] = nr_aut
    if len(nr_gon):
        nr[len(nr_aut):, :nr_gon.shape[1]] = nr_gon
    return nr


def _map_threads(fn, items, min_items=4):
    """fn over items on a thread pool (host post-processing of a batch: NumPy releases the GIL in its loops)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    items = list(items)
    if len(items) < min_items:
        return [fn(i) for i in items]
    with ThreadPoolExecutor(min(16, len(items), max(1, len(os.sched_getaffinity(0)) - 1))) as pool:
        return list(pool.map(fn, items))


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
    prepared, genders, ref_genders, n_reads = [], [], [], []
    for sample, bs in zip(samples, binsizes):
        n_reads.append(int(sum(int(np.sum(sample[x], dtype=np.int64)) for x in sample.keys())))
        sample = scale_sample(dict(sample), int(bs), int(ref_file["binsize"]))
        sample, g, rg = resolve_genders(args, sample, ref_file)
        prepared.append(sample); genders.append(g); ref_genders.append(rg)
    logging.info("Normalizing autosomes ...")
    r, z, w, n, m_lr, m_z = normalize_batch(args, prepared, ref_file, "A", eng)
    logging.info("Normalizing gonosomes ...")
    gon, nrs, rows, weights, row_of = {}, {}, {}, {}, {}
    for rg in sorted(set(ref_genders)):
        ids = [i for i, x in enumerate(ref_genders) if x == rg]
        r2, z2, w2, n2, _, _ = normalize_batch(args, [prepared[i] for i in ids], ref_file, rg, eng)
        for j, i in enumerate(ids):
            gon[i] = (r2[j], z2[j], w2, n2[j])
            row_of[i] = j
        nrs[rg] = stacked_null_ratios(ref_file, rg)
        weights[rg] = shared_weights(w, w2)
        rows[rg] = _Rows(len(ids), ref_file["mask.{}".format(rg)])

    def one(i):
        aut = (r[i], z[i], w, n[i], float(m_lr[i]), float(m_z[i]))
        rg = ref_genders[i]
        return assemble(args, aut, gon[i], nrs[rg], ref_file, rg, genders[i], n_reads[i], weights[rg], rows[rg], row_of[i])

    out = _map_threads(one, range(len(prepared)))  # a dozen NumPy passes over 2e5 bins per sample: the GIL is released in them
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
