"""Host-side mirror of the reference's newref numeric tools (newref_tools.py) on top of the
C-ABI.  Same names, argument meaning and return types as the reference:

    get_reference(pca_corrected_data, masked_bins_per_chr, masked_bins_per_chr_cum,
                  ref_size, part, split_parts) -> (int32[rows,k], float64[rows,k], float64[rows,M])

(reference newref_tools.py:155-224).  The null-sample draw uses Python's global ``random`` exactly
like the reference (``random.sample(range(S), min(S, 100))``, one draw per call, :214-217), so a
caller that seeds ``random`` gets the reference's columns.
"""
from __future__ import annotations

import ctypes
import random

import numpy as np

from . import _lib


def _get_part(partnum, outof, bincount):
    """Row range of part ``partnum`` (0-based) of ``outof`` (reference newref_tools.py:244-247)."""
    start_bin = int(bincount / float(outof) * partnum)
    end_bin = int(bincount / float(outof) * (partnum + 1))
    return start_bin, end_bin


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


class NewrefEngine:
    """Device-resident state for repeated get_reference calls on one matrix."""

    def __init__(self, device: int = 0, ctx: _lib.Context | None = None):
        self.ctx = ctx or _lib.default_context(device)
        self.n = self.s = 0

    def load(self, x, per, cum, on_device_ptr: int | None = None, shape=None, copy_from_any_device: bool = False):
        """copy_from_any_device: on_device_ptr may live on another GPU of the box; it is copied into this context once
        (wcx_newref_load x_on_device = 3) instead of being borrowed."""
        L = _lib.load()
        per = np.ascontiguousarray(per, dtype=np.int64)
        cum = np.ascontiguousarray(cum, dtype=np.int64)
        if on_device_ptr is not None:
            n, s = shape
            _lib.check(L.wcx_newref_load(self.ctx.handle, ctypes.c_void_p(on_device_ptr), n, s, _ptr(per),
                                         _ptr(cum), len(cum), 3 if copy_from_any_device else 1))
        else:
            x = np.ascontiguousarray(x, dtype=np.float64)
            n, s = x.shape
            _lib.check(L.wcx_newref_load(self.ctx.handle, _ptr(x), n, s, _ptr(per), _ptr(cum), len(cum), 0))
        self.n, self.s = int(n), int(s)

    def topk(self, row_begin, row_end, ref_size, kernel=_lib.KERNEL_AUTO, out=None, device_out=None):
        L = _lib.load()
        rows = row_end - row_begin
        if device_out is not None:
            ip, dp = device_out
            _lib.check(L.wcx_newref_topk(self.ctx.handle, row_begin, row_end, ref_size, kernel,
                                         ctypes.c_void_p(ip), ctypes.c_void_p(dp), 1))
            return None
        if out is None:
            idx = np.empty((rows, ref_size), dtype=np.int32)
            dist = np.empty((rows, ref_size), dtype=np.float64)
        else:
            idx, dist = out
        _lib.check(L.wcx_newref_topk(self.ctx.handle, row_begin, row_end, ref_size, kernel, _ptr(idx), _ptr(dist), 0))
        return idx, dist

    def null_ratios(self, row_begin, row_end, ref_size, sample_ids, idx=None, out=None, device_out=None):
        L = _lib.load()
        ids = np.ascontiguousarray(sample_ids, dtype=np.int32)
        rows = row_end - row_begin
        iptr = None if idx is None else _ptr(np.ascontiguousarray(idx, dtype=np.int32))
        if device_out is not None:
            _lib.check(L.wcx_newref_null_ratios(self.ctx.handle, iptr, 0, row_begin, row_end, ref_size, _ptr(ids),
                                                len(ids), ctypes.c_void_p(device_out), 1))
            return None
        if out is None:
            out = np.empty((rows, len(ids)), dtype=np.float64)
        _lib.check(L.wcx_newref_null_ratios(self.ctx.handle, iptr, 0, row_begin, row_end, ref_size, _ptr(ids),
                                            len(ids), _ptr(out), 0))
        return out

    def reference(self, row_begin, row_end, ref_size, sample_ids, kernel=_lib.KERNEL_AUTO, out=None, device_out=None):
        """Top-k and null ratios of the loaded matrix in one call (wcx_newref_reference: one sweep, then re-rank and
        null-ratio kernels block by block on two streams).  device_out = (idx_ptr, dist_ptr, null_ptr) leaves the
        results in HBM."""
        L = _lib.load()
        ids = np.ascontiguousarray(sample_ids, dtype=np.int32)
        rows = row_end - row_begin
        if device_out is not None:
            ip, dp, npt = device_out
            _lib.check(L.wcx_newref_reference(self.ctx.handle, row_begin, row_end, ref_size, kernel, _ptr(ids), len(ids),
                                              ctypes.c_void_p(ip), ctypes.c_void_p(dp), ctypes.c_void_p(npt), 1))
            return None
        if out is None:
            out = (np.empty((rows, ref_size), dtype=np.int32), np.empty((rows, ref_size), dtype=np.float64),
                   np.empty((rows, len(ids)), dtype=np.float64))
        idx, dist, nr = out
        _lib.check(L.wcx_newref_reference(self.ctx.handle, row_begin, row_end, ref_size, kernel, _ptr(ids), len(ids),
                                          _ptr(idx), _ptr(dist), _ptr(nr), 0))
        return idx, dist, nr

    def get_reference_host(self, x, per, cum, row_begin, row_end, ref_size, sample_ids, kernel=_lib.KERNEL_AUTO, out=None):
        """One C-ABI call (wcx_get_reference): host X in, host (indexes, distances, null ratios) out; the
        D2H copy of indexes / distances overlaps the null-ratio kernels."""
        L = _lib.load()
        x = np.ascontiguousarray(x, dtype=np.float64)
        per = np.ascontiguousarray(per, dtype=np.int64)
        cum = np.ascontiguousarray(cum, dtype=np.int64)
        ids = np.ascontiguousarray(sample_ids, dtype=np.int32)
        rows = row_end - row_begin
        if out is None:
            out = (np.empty((rows, ref_size), dtype=np.int32), np.empty((rows, ref_size), dtype=np.float64),
                   np.empty((rows, len(ids)), dtype=np.float64))
        idx, dist, nr = out
        n, s = x.shape
        _lib.check(L.wcx_get_reference(self.ctx.handle, _ptr(x), n, s, _ptr(per), _ptr(cum), len(cum), ref_size, row_begin,
                                       row_end, _ptr(ids), len(ids), kernel, _ptr(idx), _ptr(dist), _ptr(nr)))
        self.n, self.s = int(n), int(s)
        return idx, dist, nr

    def stats(self):
        out = np.zeros(8, dtype=np.int64)
        _lib.check(_lib.load().wcx_newref_stats(self.ctx.handle, _ptr(out)))
        return {"work_items": int(out[0]), "exact_fallback_rows": int(out[1]), "launches": int(out[2]),
                "rows_main_sweep": int(out[3]), "kernel": int(out[4]), "exact_compactions": int(out[5]),
                "stream_compactions": int(out[6]), "ladder_steps": int(out[7])}

    def stage_ms(self):
        out = np.zeros(8, dtype=np.float64)
        _lib.check(_lib.load().wcx_newref_stage_ms(self.ctx.handle, _ptr(out)))
        return {"sweep": out[0], "rerank": out[1], "exact_rows": out[2], "null_ratios": out[3], "prep": out[4],
                "exact_evals": out[5], "gathered_entries": out[6], "sweep_tail_past_main": out[7]}


def get_reference(pca_corrected_data, masked_bins_per_chr, masked_bins_per_chr_cum, ref_size, part,
                  split_parts, kernel=_lib.KERNEL_AUTO, device: int = 0, sample_ids=None):
    """Drop-in for the reference's get_reference (newref_tools.py:155): within-sample reference
    bins, their distances and the null ratios for part ``part`` (1-based) of ``split_parts``."""
    x = np.ascontiguousarray(pca_corrected_data, dtype=np.float64)
    cum = np.asarray(masked_bins_per_chr_cum, dtype=np.int64)
    bincount = int(cum[-1])
    start_num, end_num = _get_part(part - 1, split_parts, bincount)
    n_samples = x.shape[1]
    if sample_ids is None:
        # same draw as the reference: one random.sample per get_reference call (:214-217)
        sample_ids = random.sample(range(n_samples), min(n_samples, 100))
    eng = NewrefEngine(device)
    return eng.get_reference_host(x, masked_bins_per_chr, cum, start_num, end_num, ref_size, sample_ids, kernel)


# ---------------------------------------------------------------------------------------------
# newref preparation (reference newref_tools.py:110-147, newref_control.py:38-58)
# ---------------------------------------------------------------------------------------------
class _PCAModel:
    """What the reference keeps of the fitted sklearn PCA: ``components_`` and ``mean_``
    (newref_control.py:78-79)."""

    def __init__(self, components, mean, explained_variance):
        self.components_ = components
        self.mean_ = mean
        self.explained_variance_ = explained_variance
        self.n_components_ = components.shape[0]


_blas_controller = []


def _eigh_one_thread(gram):
    """np.linalg.eigh of the S x S Gram matrix on ONE BLAS thread.  For S <= 500 the threaded LAPACK path buys nothing
    (11 / 49 ms at S = 250 / 500 either way), but while the previous pass is still being copied out and deflated on the
    other cores its spinning worker threads make the call 10 - 30 times slower (0.24 - 0.33 s measured in the F / M
    passes of `newref` at 500 samples: most of `prep.F` / `prep.M`)."""
    import contextlib
    try:
        if not _blas_controller:
            from threadpoolctl import ThreadpoolController
            _blas_controller.append(ThreadpoolController())
        limit = _blas_controller[0].limit(limits=1, user_api="blas")
    except Exception:  # threadpoolctl comes with scikit-learn (a dependency of the reference); without it: the plain call
        limit = contextlib.nullcontext()
    with limit:
        return np.linalg.eigh(gram)


def _pca_spectrum(gram, pcacomp):
    """Top eigenpairs of the S x S Gram matrix of the centred data -> (u [S, n_eff], sigma [n_eff], lam [pcacomp],
    n_eff).  The centred matrix has rank <= S - 1: with S <= pcacomp samples (the reference allows gonosomal passes
    with exactly 5, main.py:104,119) the trailing eigenvalues are round-off, and dividing by their square root would
    blow the stored components up to ~1e130.  Only the numerically non-zero part of the spectrum (relative 1e-12) is
    sent to the device; _finish_components pads the rest."""
    w, u = _eigh_one_thread(gram)
    order = np.argsort(w)[::-1][:pcacomp]
    lam = np.clip(w[order], 0.0, None)
    lam_full = np.zeros(pcacomp)
    lam_full[: len(lam)] = lam
    n_eff = int(np.sum(lam > 1e-12 * max(lam[0], 1e-300))) if len(lam) else 0
    n_eff = max(1, min(n_eff, gram.shape[0] - 1 if gram.shape[0] > 1 else 1))
    lam_full[n_eff:] = 0.0
    return (np.ascontiguousarray(u[:, order[:n_eff]]), np.ascontiguousarray(np.sqrt(np.clip(lam[:n_eff], 1e-300, None))),
            lam_full, n_eff)


def _finish_components(comps, pcacomp):
    """sklearn's sign convention (svd_flip, u_based_decision=False: largest |entry| of each row positive) and, for a
    rank-deficient fit, unit-norm rows orthogonal to the fitted ones in place of the null-space vectors LAPACK would
    return (they reconstruct nothing of the training data; sklearn's are equally arbitrary but also unit norm)."""
    n_eff, n = comps.shape
    piv = np.argmax(np.abs(comps), axis=1)
    sgn = np.sign(comps[np.arange(n_eff), piv])
    comps = comps * np.where(sgn == 0, 1.0, sgn)[:, None]
    if n_eff < pcacomp:
        rng = np.random.default_rng(0)
        rows = [comps]
        basis = comps
        for _ in range(pcacomp - n_eff):
            v = rng.standard_normal(n)
            for _ in range(2):  # two Gram-Schmidt sweeps
                v -= basis.T @ (basis @ v)
            v /= np.linalg.norm(v)
            v *= np.sign(v[np.argmax(np.abs(v))])
            rows.append(v[None, :])
            basis = np.concatenate([basis, v[None, :]])
        comps = np.concatenate(rows)
    return np.ascontiguousarray(comps)


def stack_counts(samples, chrs):
    """int32 [bins_total, S]: per-chromosome read counts of every sample, zero padded to the longest
    sample (host half of normalize_and_mask, newref_tools.py:114-122).  One blocked transposition on host threads
    (wcx_host_stack_counts, csrc/host_newref.cu) instead of one strided column per (chromosome, sample)."""
    import os
    chrs = list(chrs)
    ns, nchr = len(samples), len(chrs)
    cols = [np.ascontiguousarray(s[str(c)], dtype=np.int32).reshape(-1) for s in samples for c in chrs]  # kept alive below
    lens = np.array([len(v) for v in cols], dtype=np.int64).reshape(ns, nchr)
    offs = np.concatenate([[0], np.cumsum(lens.max(axis=0) if ns else np.zeros(nchr, dtype=np.int64))]).astype(np.int64)
    out = np.empty((int(offs[-1]), ns), dtype=np.int32)
    if out.size:
        ptrs = np.array([v.ctypes.data for v in cols], dtype=np.uintp)
        _lib.check(_lib.load().wcx_host_stack_counts(_ptr(ptrs), _ptr(lens), ns, nchr, _ptr(offs), _ptr(out),
                                                     max(1, min(16, len(os.sched_getaffinity(0))))))
    return out


def column_totals(counts):
    """Exact (int64) read total of every sample of a stacked count matrix, row blocks on a thread pool."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    rows = counts.shape[0]
    nthreads = max(1, min(16, len(os.sched_getaffinity(0))))
    step = max(4096, -(-rows // (2 * nthreads)))
    with ThreadPoolExecutor(nthreads) as pool:
        parts = list(pool.map(lambda a: np.sum(counts[a:a + step], 0, dtype=np.int64), range(0, rows, step)))
    return np.sum(parts, 0, dtype=np.int64) if parts else np.zeros(counts.shape[1], dtype=np.int64)


def bin_sums(counts, col_sum, cols=None):
    """np.sum(counts[:, cols] / col_sum, 1) of get_mask (newref_tools.py:94-97), the same float64 values (one division
    per element, NumPy's pairwise order along the row), on host threads and without the float matrix."""
    import os
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    col_sum = np.ascontiguousarray(col_sum, dtype=np.float64)
    cols = None if cols is None else np.ascontiguousarray(cols, dtype=np.int32)
    assert len(col_sum) == (counts.shape[1] if cols is None else len(cols))
    out = np.empty(counts.shape[0], dtype=np.float64)
    _lib.check(_lib.load().wcx_host_bin_sums(_ptr(counts), counts.shape[0], counts.shape[1], None if cols is None else _ptr(cols),
                                             0 if cols is None else len(cols), _ptr(col_sum), _ptr(out),
                                             max(1, min(16, len(os.sched_getaffinity(0))))))
    return out


def take_columns(counts, rows, cols):
    """counts[:rows][:, cols] as a new C-contiguous matrix (the count matrix of a gonosomal pass: a row prefix and a
    column subset of the stacked matrix of all samples), row blocks on a thread pool: the plain fancy index is one
    strided gather on one core (0.8 s for 250 of 500 columns at 15 kb, 0.05 s here)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    cols = np.ascontiguousarray(cols, dtype=np.int64)
    rows = int(rows)
    out = np.empty((rows, len(cols)), dtype=counts.dtype)
    nthreads = max(1, min(16, len(os.sched_getaffinity(0))))
    step = max(1024, -(-rows // (4 * nthreads)))

    def block(a):
        np.take(counts[a:min(rows, a + step)], cols, axis=1, out=out[a:min(rows, a + step)])

    with ThreadPoolExecutor(nthreads) as pool:
        list(pool.map(block, range(0, rows, step)))
    return out


def normalize_and_mask(samples, chrs, mask, device: int = 0):
    """Drop-in for newref_tools.normalize_and_mask (reference :110-129): read-depth normalisation
    (each sample divided by its total) and masking.  Bit-exact with the reference."""
    counts = np.ascontiguousarray(stack_counts(samples, chrs))
    pos = np.ascontiguousarray(np.flatnonzero(np.asarray(mask, dtype=bool)[: counts.shape[0]]), dtype=np.int32)
    out = np.empty((len(pos), counts.shape[1]), dtype=np.float64)
    ctx = _lib.default_context(device)
    ctx.__dict__["_counts_obj"] = None  # this call replaces whatever count matrix DevicePrep left on the device
    _lib.check(_lib.load().wcx_newref_normalize_and_mask(ctx.handle, _ptr(counts), counts.shape[0], counts.shape[1], _ptr(pos),
                                                         len(pos), _ptr(out), 0))
    return out


def train_pca(ref_data, pcacomp=5, device: int = 0, keep_on_device: bool = False):
    """Drop-in for newref_tools.train_pca (reference :138-147): returns (corrected [N, S], pca) with
    pca.components_ [pcacomp, N] and pca.mean_ [N].  The model is the exact PCA that sklearn's
    randomized solver approximates: Gram matrix and correction on the GPU, the
    S x S eigen-decomposition on the host."""
    x = np.ascontiguousarray(ref_data, dtype=np.float64)
    n, s = x.shape
    L = _lib.load()
    ctx = _lib.default_context(device)
    mean = np.empty(n, dtype=np.float64)
    gram = np.empty((s, s), dtype=np.float64)
    _lib.check(L.wcx_pca_gram(ctx.handle, _ptr(x), n, s, 0, _ptr(mean), _ptr(gram)))
    u, sigma, lam, n_eff = _pca_spectrum(gram, pcacomp)
    comps = np.empty((n_eff, n), dtype=np.float64)
    corrected = None if keep_on_device else np.empty((n, s), dtype=np.float64)
    _lib.check(L.wcx_pca_apply(ctx.handle, _ptr(u), _ptr(sigma), n_eff, _ptr(comps),
                               None if keep_on_device else _ptr(corrected), 0))
    return corrected, _PCAModel(_finish_components(comps, pcacomp), mean, lam / max(s - 1, 1))


def pca_distance(corrected=None, shape=None, device: int = 0):
    """Per-sample median profile and per-bin squared distance to it (newref_control.py:40-41).
    corrected=None uses the matrix left on the device by train_pca(keep_on_device=True)."""
    L = _lib.load()
    ctx = _lib.default_context(device)
    if corrected is None:
        n, s = shape
        med = np.empty(s); d = np.empty(n)
        _lib.check(L.wcx_pca_distance(ctx.handle, None, n, s, 1, _ptr(med), _ptr(d)))
    else:
        x = np.ascontiguousarray(corrected, dtype=np.float64)
        n, s = x.shape
        med = np.empty(s); d = np.empty(n)
        _lib.check(L.wcx_pca_distance(ctx.handle, _ptr(x), n, s, 0, _ptr(med), _ptr(d)))
    return d, med


# ---------------------------------------------------------------------------------------------
# device-resident form of the preparation chain (no [N, S] matrix crosses PCIe between the steps)
# ---------------------------------------------------------------------------------------------
class DevicePrep:
    """normalize_and_mask -> train_pca -> PCA-distance -> get_reference with the [N, S] matrices kept in the
    context's device buffers: the same C-ABI calls as the drop-in functions above with NULL host pointers
    (`out = NULL`, `x = NULL`, `corrected = NULL`, `wcx_newref_load(x_on_device = 2)`)."""

    def __init__(self, device: int = 0):
        self.device = device
        self.ctx = _lib.default_context(device)
        self.shape = None

    def normalize_and_mask(self, counts, mask):
        """counts: int32 [bins_total, S] (stack_counts); mask: bool [bins_total].  Leaves the [N, S] matrix on the device."""
        counts = np.ascontiguousarray(counts)
        pos = np.ascontiguousarray(np.flatnonzero(np.asarray(mask, dtype=bool)[: counts.shape[0]]), dtype=np.int32)
        # the same matrix again (the redo after the PCA-distance filter): it is still on the device
        resident = self.ctx.__dict__.get("_counts_obj") is counts  # residency is a property of the context
        self.ctx.__dict__["_counts_obj"] = None
        _lib.check(_lib.load().wcx_newref_normalize_and_mask(self.ctx.handle, None if resident else _ptr(counts), counts.shape[0],
                                                             counts.shape[1], _ptr(pos), len(pos), None, 1))
        self.ctx.__dict__["_counts_obj"] = counts
        self.shape = (len(pos), counts.shape[1])
        return self.shape

    def train_pca(self, pcacomp=5):
        """PCA of the resident matrix; returns the model (components_, mean_), keeps `corrected` on the device."""
        n, s = self.shape
        L = _lib.load()
        mean = np.empty(n, dtype=np.float64)
        gram = np.empty((s, s), dtype=np.float64)
        _lib.check(L.wcx_pca_gram(self.ctx.handle, None, n, s, 1, _ptr(mean), _ptr(gram)))
        u, sigma, lam, n_eff = _pca_spectrum(gram, pcacomp)
        comps = np.empty((n_eff, n), dtype=np.float64)
        _lib.check(L.wcx_pca_apply(self.ctx.handle, _ptr(u), _ptr(sigma), n_eff, _ptr(comps), None, 0))
        return _PCAModel(_finish_components(comps, pcacomp), mean, lam / max(s - 1, 1))

    def pca_distance(self):
        return pca_distance(None, shape=self.shape, device=self.device)

    def load_into(self, engine, per, cum):
        """wcx_newref_load on the resident corrected matrix (no copy)."""
        n, s = self.shape
        per = np.ascontiguousarray(per, dtype=np.int64)
        cum = np.ascontiguousarray(cum, dtype=np.int64)
        _lib.check(_lib.load().wcx_newref_load(engine.ctx.handle, None, n, s, _ptr(per), _ptr(cum), len(cum), 2))
        engine.n, engine.s = int(n), int(s)

    def device_ptr(self, which: str = "corrected") -> int:
        """Device address of a resident matrix (for NewrefEngine.load(..., copy_from_any_device=True) on other GPUs)."""
        n, s = self.shape
        out = ctypes.c_void_p()
        _lib.check(_lib.load().wcx_prep_device_ptr(self.ctx.handle, 0 if which == "masked" else 1, n, s, ctypes.byref(out)))
        return int(out.value)

    def fetch(self, which: str):
        """Host copy of a resident matrix ("masked" or "corrected"): tests, or callers that want the reference's
        prepdatafile (newref_control.py:68)."""
        n, s = self.shape
        out = np.empty((n, s), dtype=np.float64)
        _lib.check(_lib.load().wcx_prep_fetch(self.ctx.handle, 0 if which == "masked" else 1, n, s, _ptr(out)))
        return out
