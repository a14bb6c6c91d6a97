"""newref stage controllers (mirror of the reference's newref_control.py) on the GPU kernels.

The reference fans `get_reference` out over `cpus` parts through temporary .npz part files
(newref_control.py:90-150); here the parts are computed by the GPU(s) and concatenated in part order
in memory, which yields the same arrays (`tool_newref_post`, :159-189).  Output keys, dtypes and
suffix rules of the final reference file follow `tool_newref_merge` (:220-237)."""
from __future__ import annotations

import logging
import random

import numpy as np

from . import _lib, newref_tools, npz_io


def tool_newref_prep(samples, gender, mask, bins_per_chr, device: int = 0, counts=None):
    """Numeric body of the reference's tool_newref_prep (newref_control.py:24-80): masking,
    normalisation, PCA correction and the PCA-distance bin filter.  `mask` is edited IN PLACE like in
    the reference (:51-54; the edit leaks into the later F / M passes).
    Returns a dict with the prep arrays and pca_corrected_data.  counts: the stacked int32 count matrix of exactly these
    samples and chromosomes (newref_tools.stack_counts) when the caller already has it."""
    last_chr = {"A": 22, "F": 23}.get(gender, 24)
    bins_per_chr = list(bins_per_chr[:last_chr])
    mask = mask[: int(np.sum(bins_per_chr))]  # a view of the caller's total mask
    chrs = range(1, last_chr + 1)
    # the [N, S] matrices stay in HBM from here to get_reference (newref_tools.DevicePrep)
    dp = newref_tools.DevicePrep(device)
    if counts is None or counts.shape != (int(np.sum(bins_per_chr)), len(samples)):
        counts = newref_tools.stack_counts(samples, chrs)
    dp.normalize_and_mask(counts, mask)
    pca = dp.train_pca()
    d, _ = dp.pca_distance()
    mad = np.median(np.abs(d - np.median(d)))
    cutoff = max(np.median(d) + 10 * mad, 5.0)  # newref_control.py:45
    bad = d > cutoff
    if np.any(bad):
        logging.info("Removing {} anomalous bins based on PCA distance (cutoff={:.4f})".format(int(np.sum(bad)), cutoff))
        masked_indices = np.where(mask)[0]
        mask[masked_indices[bad]] = False
        dp.normalize_and_mask(counts, mask)
        pca = dp.train_pca()
    corrected = dp  # device-resident: tool_newref_main loads it without a copy
    offs = np.concatenate([[0], np.cumsum(bins_per_chr)]).astype(int)
    masked_bins_per_chr = [int(np.sum(mask[offs[i]:offs[i + 1]])) for i in range(len(bins_per_chr))]
    return {
        "gender": gender,
        "mask": mask.copy(),
        "bins_per_chr": np.array(bins_per_chr),
        "masked_bins_per_chr": np.array(masked_bins_per_chr),
        "masked_bins_per_chr_cum": np.cumsum(masked_bins_per_chr),
        "pca_components": pca.components_,
        "pca_mean": pca.mean_,
        "pca_corrected_data": corrected,
    }


def tool_newref_main(prep, refsize, parts: int = 1, device: int = 0, devices=None):
    """get_reference over `parts` parts, concatenated in part order (reference tool_newref_main +
    tool_newref_post, newref_control.py:90-189).  One null-sample draw per part, like the reference.

    devices: CUDA devices to spread the target bins over (`newref --gpus N`).  The reference's own fan-out is over
    parts of the target-bin axis (newref_control.py:90-104, newref_tools.py:244-247) -- rows are independent given the
    matrix -- so every part is cut into one row range per device; each device gets the matrix once (a copy of the
    matrix prepared on `device`, over NVLink) and writes its rows straight into the result arrays.  One host thread
    drives each device through its own context (the C-ABI calls release the GIL); no collective is needed."""
    x = prep["pca_corrected_data"]
    per, cum = prep["masked_bins_per_chr"], prep["masked_bins_per_chr_cum"]
    n, s = x.shape
    m = min(s, 100)
    draws = []
    for part in range(1, parts + 1):
        start, end = newref_tools._get_part(part - 1, parts, n)
        draws.append((start, end, random.sample(range(s), m)))  # newref_tools.py:214-217, in part order
    idx = _lib.pinned.empty((n, refsize), np.int32)
    dist = _lib.pinned.empty((n, refsize), np.float64)
    nr = _lib.pinned.empty((n, m), np.float64)
    devices = list(devices) if devices else [device]

    def load(eng, first):
        if isinstance(x, np.ndarray):
            eng.load(x, per, cum)
        elif first:  # newref_tools.DevicePrep: the corrected matrix is already in this context
            x.load_into(eng, per, cum)
        else:
            eng.load(None, per, cum, on_device_ptr=x.device_ptr("corrected"), shape=(n, s), copy_from_any_device=True)

    def run(eng, g, ngpu):
        for start, end, ids in draws:
            a, b = start + (end - start) * g // ngpu, start + (end - start) * (g + 1) // ngpu
            if b > a:
                eng.reference(a, b, refsize, ids, out=(idx[a:b], dist[a:b], nr[a:b]))

    if len(devices) == 1:
        eng = newref_tools.NewrefEngine(devices[0])
        load(eng, True)
        run(eng, 0, 1)
    else:
        from concurrent.futures import ThreadPoolExecutor
        engines = [newref_tools.NewrefEngine(d, newref_tools._lib.default_context(d) if i == 0 and d == device else _lib.Context(d))
                   for i, d in enumerate(devices)]

        def worker(g):
            load(engines[g], g == 0 and devices[0] == device)
            run(engines[g], g, len(devices))

        with ThreadPoolExecutor(len(devices)) as pool:
            list(pool.map(worker, range(len(devices))))
        for e in engines[1:]:
            e.ctx.close()
    out = {k: v for k, v in prep.items() if k != "pca_corrected_data"}
    out["indexes"], out["distances"], out["null_ratios"] = idx, dist, nr
    return out


RESULT_KEYS = ("mask", "bins_per_chr", "masked_bins_per_chr", "masked_bins_per_chr_cum", "pca_components", "pca_mean",
               "indexes", "distances", "null_ratios")


def tool_newref_merge(outfile, results, binsize, is_nipt, trained_cutoff, writer=None):
    """Final reference .npz with the reference's key layout (newref_control.py:220-237): autosomal keys
    plain, gonosomal keys suffixed .F / .M, plus has_female / has_male / is_nipt / trained_cutoff.
    Same format as np.savez_compressed (:237), deflated on all cores; with `writer` (npz_io.AsyncNpzWriter) the
    arrays of the passes already handed to writer_add_pass are being deflated since their pass finished."""
    own = writer is None
    if own:
        writer = npz_io.AsyncNpzWriter(outfile)
    final_ref = {"has_female": False, "has_male": False}
    for res in results:
        gender = res["gender"]
        sfx = "" if gender == "A" else ".{}".format(gender)
        if gender == "F":
            final_ref["has_female"] = True
        if gender == "M":
            final_ref["has_male"] = True
        final_ref["binsize" + sfx] = binsize
        for key in RESULT_KEYS:
            final_ref[key + sfx] = res[key]
        if not res.get("_queued"):
            writer_add_pass(writer, res, binsize)
    for key in ("has_female", "has_male"):
        writer.add(key, final_ref[key])
    final_ref["is_nipt"] = is_nipt
    final_ref["trained_cutoff"] = trained_cutoff
    writer.add("is_nipt", is_nipt)
    writer.add("trained_cutoff", trained_cutoff)
    writer.close()
    return final_ref


def writer_add_pass(writer, res, binsize):
    """Starts deflating the arrays of one finished pass (A / F / M) while the next pass runs on the GPU."""
    sfx = "" if res["gender"] == "A" else ".{}".format(res["gender"])
    writer.add("binsize" + sfx, binsize)
    for key in RESULT_KEYS:
        writer.add(key + sfx, res[key])
    res["_queued"] = True
