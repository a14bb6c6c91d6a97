"""Host-side mirror of the reference's predict numeric tools (predict_tools.py, overall_tools.py)
on top of the C-ABI.  Function names and return values follow the reference; the arithmetic runs in
the CUDA kernels of csrc/predict.cu.

    coverage_normalize_and_mask + project_pc + normalize_repeat  -> PredictEngine.normalize_set
    get_weights(ref_file, ap)                                    (predict_tools.py:152-155)
    get_optimal_cutoff(ref_file, repeats)                        (predict_tools.py:74-82)
    get_z_score(results_c, results)                              (overall_tools.py:88-119)
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib

SET_ID = {"": 0, ".F": 1, ".M": 2}


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def map_threads(fn, items, min_items=4):
    """fn over items on a thread pool (host post-processing of a batch: NumPy releases the GIL in its loops)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    items = list(items)
    if len(items) < min_items:
        return [fn(i) for i in items]
    # (more threads than this lose to the interpreter lock between the NumPy calls: 8 cores, batch of 96: 1.26 / 0.87 /
    # 1.25 s on 2 / 4 / 6 threads)
    with ThreadPoolExecutor(min(6, len(items), max(1, len(os.sched_getaffinity(0)) // 2))) as pool:
        return list(pool.map(fn, items))


def flatten(arrays, dtype=np.float64):
    """The arrays of a list (the per-chromosome vectors of a sample, of several samples ...) as ONE contiguous vector.
    No copy when they already lie back to back in one allocation (the result assembly of this package hands out
    per-chromosome views of one row per sample, the rows of a batch being rows of one matrix): the vector is then a view
    of that allocation, and writes through the per-chromosome views (blacklisting) show in it."""
    arrays = list(arrays)
    dtype = np.dtype(dtype)
    if not arrays:
        return np.zeros(0, dtype=dtype)
    first = arrays[0]
    if len(arrays) == 1 and isinstance(first, np.ndarray) and first.dtype == dtype and first.ndim == 1 and first.flags.c_contiguous:
        return first
    owner = first.base if isinstance(first, np.ndarray) else None
    if isinstance(owner, np.ndarray) and len(arrays) > 1:
        nxt, total = first.ctypes.data if first.size else None, 0
        for a in arrays:
            if not (isinstance(a, np.ndarray) and a.base is owner and a.dtype == dtype and a.ndim == 1
                    and (a.size == 0 or (a.strides[0] == dtype.itemsize and (nxt is None or a.ctypes.data == nxt)))):
                break
            if a.size:
                nxt = a.ctypes.data + a.nbytes
                total += a.size
        else:
            if total == 0:
                return np.zeros(0, dtype=dtype)
            start = next(a for a in arrays if a.size)
            return np.lib.stride_tricks.as_strided(start, shape=(total,), strides=(dtype.itemsize,))
    return np.concatenate([np.asarray(a, dtype=dtype) for a in arrays])


def raw_vector(sample, bins_per_chr, out=None):
    """Per-chromosome read counts padded / truncated to the reference's bins_per_chr and
    concatenated (the host half of coverage_normalize_and_mask, predict_tools.py:35-44); the
    division by the total and the masking happen on the device.  `out`: row of a preallocated batch matrix."""
    total = int(np.sum(bins_per_chr))
    if out is None:
        out = np.empty(total, dtype=np.float64)
    parts = [np.asarray(sample[str(c + 1)]) for c in range(len(bins_per_chr))]
    if all(len(a) == int(nb) and a.ndim == 1 for a, nb in zip(parts, bins_per_chr)):
        np.concatenate(parts, out=out[:total], casting="unsafe")  # the usual case: sample and reference binned alike
        return out
    out[:] = 0.0
    off = 0
    for a, nb in zip(parts, bins_per_chr):
        nb = int(nb)
        m = min(nb, len(a))
        out[off:off + m] = a[:m]
        off += nb
    return out


class PredictEngine:
    """Keeps the reference arrays of one reference .npz resident on the device between calls
    (the reference re-inflates them from the .npz on every access)."""

    def __init__(self, device: int = 0, ctx: _lib.Context | None = None):
        self.ctx = ctx or _lib.default_context(device)
        # which reference object fills which device slot is a property of the CONTEXT (engines may share one), and the
        # cache holds a strong reference to the dict: an id() can be reused once the old dict is collected
        if not hasattr(self.ctx, "predict_sets"):
            self.ctx.predict_sets = {}

    @property
    def meta(self):
        return {ap: m for ap, (_, m) in self.ctx.predict_sets.items()}

    def _ensure_ref(self, ref_file, ap: str):
        sets = self.ctx.predict_sets
        if any(obj is not ref_file for obj, _ in sets.values()):
            sets.clear()  # another reference: every slot is stale
        if ap in sets:
            return
        L = _lib.load()
        idx = np.ascontiguousarray(ref_file["indexes" + ap], dtype=np.int32)
        dist = np.ascontiguousarray(ref_file["distances" + ap], dtype=np.float64)
        per = np.ascontiguousarray(ref_file["masked_bins_per_chr" + ap], dtype=np.int64)
        cum = np.ascontiguousarray(ref_file["masked_bins_per_chr_cum" + ap], dtype=np.int64)
        comps = np.ascontiguousarray(ref_file["pca_components" + ap], dtype=np.float64)
        mean = np.ascontiguousarray(ref_file["pca_mean" + ap], dtype=np.float64)
        mask = np.asarray(ref_file["mask" + ap], dtype=bool)
        bpc = np.asarray(ref_file["bins_per_chr" + ap])
        mask_pos = np.ascontiguousarray(np.flatnonzero(mask), dtype=np.int32)
        n, k = idx.shape
        _lib.check(L.wcx_predict_load_ref(self.ctx.handle, SET_ID[ap], _ptr(idx), _ptr(dist), n, k, _ptr(per), _ptr(cum),
                                          len(cum), _ptr(comps), _ptr(mean), comps.shape[0], _ptr(mask_pos), int(len(mask))))
        sets[ap] = (ref_file, {"n": n, "k": k, "cum": cum, "bins_per_chr": bpc, "bins_total": int(len(mask))})

    def stacked_null_ratios(self, ref_file, ref_gender):
        """predict_control.stacked_null_ratios, built once per resident reference (165 MB at 15 kb); the same array
        object on every call also keeps the copy on the device valid (segment_zscore uploads per array object)."""
        self._ensure_ref(ref_file, "")
        cache = self.ctx.predict_sets[""][1].setdefault("stacked_nr", {})
        if ref_gender not in cache:
            from .predict_control import stacked_null_ratios
            cache[ref_gender] = stacked_null_ratios(ref_file, ref_gender)
        return cache[ref_gender]

    def get_weights(self, ref_file, ap):
        """get_weights (predict_tools.py:152-155); a function of the reference only: computed once per resident set."""
        self._ensure_ref(ref_file, ap)
        meta = self.ctx.predict_sets[ap][1]
        if "weights" not in meta:
            out = np.empty(meta["n"], dtype=np.float64)
            _lib.check(_lib.load().wcx_predict_weights(self.ctx.handle, SET_ID[ap], _ptr(out)))
            meta["weights"] = out
        return meta["weights"].copy()

    def get_optimal_cutoff(self, ref_file, repeats):
        """get_optimal_cutoff (predict_tools.py:74-82), always on the autosomal distances (:75); ten streaming passes
        over 0.46 GB at 15 kb, cached per (resident reference, repeats)."""
        self._ensure_ref(ref_file, "")
        cache = self.ctx.predict_sets[""][1].setdefault("cutoff", {})
        if int(repeats) not in cache:
            out = ctypes.c_double()
            _lib.check(_lib.load().wcx_predict_optimal_cutoff(self.ctx.handle, 0, int(repeats), ctypes.byref(out)))
            cache[int(repeats)] = float(out.value)
        return cache[int(repeats)]

    def normalize_set(self, samples, ref_file, ap, cutoff, cp, ct):
        """Batch form: samples = list of sample dicts -> (z, r, nref [B, n - ct], m_lr [B], m_z [B])."""
        self._ensure_ref(ref_file, ap)
        meta = self.meta[ap]
        # page-locked staging for everything that crosses PCIe (0.6 GB at batch 96); _lib.pinned recycles the buffers
        b = len(samples)
        # (a handful of samples: ordinary arrays -- cudaHostAlloc costs more than the pageable copy of a few MB)
        alloc = _lib.pinned.empty if b >= 8 else (lambda shape: np.empty(shape, dtype=np.float64))
        raw = alloc((b, meta["bins_total"]))
        map_threads(lambda i: raw_vector(samples[i], meta["bins_per_chr"], out=raw[i]), range(b))
        nout = meta["n"] - ct
        z = alloc((b, nout)); r = alloc((b, nout)); nref = alloc((b, nout))
        m_lr = np.empty(b); m_z = np.empty(b)
        _lib.check(_lib.load().wcx_predict_normalize(self.ctx.handle, SET_ID[ap], _ptr(raw), b, float(cutoff), int(cp), int(ct),
                                                     _ptr(z), _ptr(r), _ptr(nref), _ptr(m_lr), _ptr(m_z)))
        accumulate_ms(self.ctx, ("coverage_project", "gather_list", "passes", "medians"))
        return z, r, nref, m_lr, m_z

    def stage_ms(self):
        """Device milliseconds of the last calls (wcx_predict_stage_ms)."""
        return stage_ms(self.ctx)

    def segment_zscore(self, nr, inflate_pos, r, w, seg_se, seg_r):
        nr = np.ascontiguousarray(nr, dtype=np.float64)
        inflate_pos = np.ascontiguousarray(inflate_pos, dtype=np.int32)
        r = np.ascontiguousarray(r, dtype=np.float64)
        w = np.ascontiguousarray(w, dtype=np.float64)
        seg_se = np.ascontiguousarray(seg_se, dtype=np.int64).reshape(-1, 2)
        seg_r = np.ascontiguousarray(seg_r, dtype=np.float64)
        out = np.empty(len(seg_r), dtype=np.float64)
        # the null ratios belong to the reference: upload them once per array object (strong reference kept on the
        # context), later calls pass NULL
        resident = getattr(self.ctx, "z_nr_obj", None) is nr
        self.ctx.z_nr_obj = None
        _lib.check(_lib.load().wcx_segment_zscore(self.ctx.handle, None if resident else _ptr(nr), nr.shape[0], nr.shape[1],
                                                  _ptr(inflate_pos), _ptr(r), _ptr(w), len(r), _ptr(seg_se), _ptr(seg_r),
                                                  len(seg_r), _ptr(out)))
        if len(seg_r):
            self.ctx.z_nr_obj = nr
            accumulate_ms(self.ctx, ("segment_z",))
        return out


def stage_ms(ctx):
    out = np.zeros(8)
    _lib.check(_lib.load().wcx_predict_stage_ms(ctx.handle, _ptr(out)))
    return {"coverage_project": out[0], "normalize_repeat": out[1], "segment_z": out[2], "cbs": out[3],
            "gather_list": out[4], "passes": out[5], "medians": out[6]}


def accumulate_ms(ctx, keys):
    """Adds the device times of the call that just returned to the context's running totals (bench / CLI timings)."""
    acc = ctx.__dict__.setdefault("kernel_ms_acc", {})
    sm = stage_ms(ctx)
    for k in keys:
        acc[k] = acc.get(k, 0.0) + float(sm[k])


_engines = {}


def default_engine(device: int = 0) -> PredictEngine:
    if device not in _engines:
        _engines[device] = PredictEngine(device)
    return _engines[device]


def get_weights(ref_file, ap, engine: PredictEngine | None = None):
    """Drop-in for predict_tools.get_weights (reference predict_tools.py:152-155)."""
    return (engine or default_engine()).get_weights(ref_file, ap)


def get_optimal_cutoff(ref_file, repeats, engine: PredictEngine | None = None):
    """Drop-in for predict_tools.get_optimal_cutoff (reference predict_tools.py:74-82)."""
    return (engine or default_engine()).get_optimal_cutoff(ref_file, repeats)


def get_z_score_batch(items, engine: PredictEngine | None = None):
    """get_z_score for several samples that share ONE null-ratio array (same reference gender): items =
    [(results_c, results), ...] with results["results_nr"] = {"dense": nr, "inflate": ...} (the array form).  The bin
    axes of the samples are concatenated and the segment offsets shifted, so one device call serves the batch.
    Returns a list of per-sample lists like get_z_score."""
    if not items:
        return []
    nr = items[0][1]["results_nr"]["dense"]
    segs, segr, counts = [], [], []
    base = 0
    for results_c, results in items:
        assert results["results_nr"]["dense"] is nr
        offs = np.concatenate([[0], np.cumsum([len(x) for x in results["results_r"]])]).astype(np.int64)
        sg = np.array([[s[0], s[1], s[2]] for s in results_c], dtype=np.int64).reshape(-1, 3)
        segs.append(np.stack([base + offs[sg[:, 0]] + sg[:, 1], base + offs[sg[:, 0]] + sg[:, 2]], axis=1))
        segr.append(np.array([s[3] for s in results_c], dtype=np.float64))
        counts.append(len(results_c))
        base += int(offs[-1])
    # (no copies when the samples come from predict_control.assemble_batch: their rows are rows of one matrix)
    r = flatten([x for _, res in items for x in res["results_r"]])
    w = flatten([x for _, res in items for x in res["results_w"]])
    infl = flatten([res["results_nr"]["inflate"] for _, res in items], np.int32)
    z = (engine or default_engine()).segment_zscore(nr, infl, r, w, np.concatenate(segs), np.concatenate(segr))
    out, o = [], 0
    for c in counts:
        out.append([("nan" if np.isnan(v) else float(v)) for v in z[o:o + c]])
        o += c
    return out


def get_z_score(results_c, results, engine: PredictEngine | None = None):
    """Drop-in for overall_tools.get_z_score (reference overall_tools.py:88-119): takes the
    reference's per-chromosome list structures and returns a list of floats / the string "nan"."""
    results_nr, results_r, results_w = results["results_nr"], results["results_r"], results["results_w"]
    r = flatten(results_r)
    w = flatten(results_w)
    offs = np.concatenate([[0], np.cumsum([len(x) for x in results_r])]).astype(np.int64)
    if isinstance(results_nr, dict):
        # array form used by this package's own pipeline: dense [n_masked, M] rows + unmasked-bin -> row map
        nr, inflate = results_nr["dense"], results_nr["inflate"]
    else:
        rows = []
        inflate = np.full(len(r), -1, dtype=np.int32)
        g = 0
        for chrom in results_nr:
            for row in chrom:
                if not isinstance(row, (int, float)):
                    inflate[g] = len(rows)
                    rows.append(row)
                g += 1
        if not rows:
            return ["nan" for _ in results_c]
        width = max(len(x) for x in rows)
        nr = np.full((len(rows), width), np.nan)
        for i, x in enumerate(rows):
            nr[i, :len(x)] = x
    seg_se = np.array([[offs[s[0]] + s[1], offs[s[0]] + s[2]] for s in results_c], dtype=np.int64).reshape(-1, 2)
    seg_r = np.array([s[3] for s in results_c], dtype=np.float64)
    z = (engine or default_engine()).segment_zscore(nr, inflate, r, w, seg_se, seg_r)
    return [("nan" if np.isnan(v) else float(v)) for v in z]
