"""A stand-in for libwcx_b200.so for HOST-LOGIC tests and host-side profiling without a GPU (test infrastructure: only
tests/ and tools/host_profile_mock.py import it; nothing under wisecondorx_b200/ does, and there is no way to select it
from the package -- a test installs it by assigning `wisecondorx_b200._lib._lib`).

The entry points the predict flow calls fill their outputs with cheap synthetic values of the right shape (ratios
around 1 with exact zeros / NaN sprinkled in, z-scores around 0, one to three segments per chromosome).  THE NUMBERS MEAN
NOTHING; what the tests check is that the Python side around the device calls (batching, result assembly, CBS.R pre- and
post-processing, z-score plumbing) gives every sample of a batch exactly what it gets alone, and what the profile
measures is the time that side takes.  Signatures follow include/wcx_b200.h."""
import ctypes
import time

import numpy as np

from wisecondorx_b200 import synth


def _arr(ptr, shape, dtype):
    n = int(np.prod(shape))
    addr = ptr.value if isinstance(ptr, ctypes.c_void_p) else int(ptr)
    buf = (ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(addr)
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)


class FakeLib:
    """See the module docstring."""

    def __init__(self, nasty=False, real=None):
        self.nasty = nasty
        if real is None:
            from wisecondorx_b200 import _lib
            saved, _lib._lib = _lib._lib, None
            try:
                real = _lib.load()
            finally:
                _lib._lib = saved
        # host-only entry points (no device work) are the product's own: the real library serves them
        self.real = real
        self.wcx_predict_assemble = real.wcx_predict_assemble
        self.sets = {}
        self.t = 0.0
        self.rng = np.random.default_rng(0)
        self.noise = self.rng.standard_normal(1 << 22)

    def __getattr__(self, name):
        if name in ("wcx_host_stack_counts", "wcx_host_bin_sums", "wcx_cbs_pack_count", "wcx_cbs_pack", "wcx_cbs_unpack",
                    "wcx_host_format_bins", "wcx_host_format_repr"):  # host-only: the real library
            return getattr(self.real, name)
        raise AttributeError(name)

    def wcx_last_error(self):
        return self.real.wcx_last_error()

    def wcx_create(self, device, ref):
        ref._obj.value = 1
        return 0

    def wcx_destroy(self, h):
        return None

    def wcx_host_alloc(self, size, ref):
        buf = np.empty(int(size), dtype=np.uint8)
        self.sets.setdefault("_keep", []).append(buf)
        ref._obj.value = buf.ctypes.data
        return 0

    def wcx_host_free(self, p):
        return 0

    # ---- newref: shapes only (ones, noise, a few far-off bins so that the PCA-distance filter and its redo run) ----------
    def wcx_newref_normalize_and_mask(self, h, counts, rows, s, pos, npos, out, keep):
        self.prep_shape = (int(npos), int(s))
        return 0

    def wcx_pca_gram(self, h, x, n, s, on_device, mean, gram):
        _arr(mean, (n,), np.float64)[:] = 1.0 / max(n, 1)
        a = self.noise[:s * (s + 3)].reshape(s, s + 3)
        _arr(gram, (s, s), np.float64)[:] = a @ a.T
        return 0

    def wcx_pca_apply(self, h, u, sigma, n_eff, comps, corrected, flag):
        n = self.prep_shape[0]
        _arr(comps, (n_eff, n), np.float64)[:] = self.noise[:n_eff * n].reshape(n_eff, n)
        return 0

    def wcx_pca_distance(self, h, x, n, s, on_device, med, d):
        _arr(med, (s,), np.float64)[:] = 1.0
        dd = _arr(d, (n,), np.float64)
        dd[:] = 1.0 + 0.01 * self.noise[:n]
        dd[5::97] = 100.0
        return 0

    def wcx_newref_load(self, h, x, n, s, per, cum, nchr, mode):
        self.loaded = (int(n), int(s))
        return 0

    def wcx_newref_reference(self, h, a, b, refsize, kernel, ids, nids, idx, dist, nr, on_device):
        rows = int(b - a)
        _arr(idx, (rows, refsize), np.int32)[:] = np.arange(refsize, dtype=np.int32)
        _arr(dist, (rows, refsize), np.float64)[:] = 1.0 + np.arange(refsize) * 0.01
        # a function of the absolute row, so that any split of the rows over calls / devices gives the same arrays
        k = (np.arange(int(a), int(b))[:, None] * nids + np.arange(nids)[None, :]) % len(self.noise)
        _arr(nr, (rows, nids), np.float64)[:] = 0.03 * self.noise[k]
        return 0

    def wcx_prep_device_ptr(self, h, which, n, s, ref):
        ref._obj.value = 4096
        return 0

    def wcx_newref_stats(self, h, out):
        return 0

    def wcx_newref_stage_ms(self, h, out):
        return 0

    def wcx_newref_prep_stage_ms(self, h, out):
        return 0

    def wcx_sync(self, h):
        return 0

    def wcx_predict_load_ref(self, h, sid, idx, dist, n, k, per, cum, nchr, comps, mean, ncomp, mask_pos, bins_total):
        self.sets[sid] = (int(n), int(k), int(bins_total))
        return 0

    def wcx_predict_weights(self, h, sid, out):
        t0 = time.perf_counter()
        _arr(out, (self.sets[sid][0],), np.float64)[:] = 1.0
        self.t += time.perf_counter() - t0
        return 0

    def wcx_predict_optimal_cutoff(self, h, sid, repeats, ref):
        ref._obj.value = 3.0
        return 0

    def wcx_predict_normalize(self, h, sid, raw, b, cutoff, cp, ct, z, r, nref, m_lr, m_z):
        t0 = time.perf_counter()
        n, bins_total = self.sets[sid][0], self.sets[sid][2]
        nout = n - int(ct)
        rawm = _arr(raw, (b, bins_total), np.float64)
        zz, rr, nn = _arr(z, (b, nout), np.float64), _arr(r, (b, nout), np.float64), _arr(nref, (b, nout), np.float64)
        ml, mz = _arr(m_lr, (b,), np.float64), _arr(m_z, (b,), np.float64)
        for i in range(b):
            # a function of the sample's own counts, not of its position in the batch
            key = int(rawm[i, ::97].sum()) + 131 * sid
            o = (key * 7919) % (len(self.noise) - nout - 1)
            zz[i] = self.noise[o:o + nout]
            rr[i] = 1.0 + 0.05 * self.noise[o + 1:o + 1 + nout]
            nn[i] = 300.0
            if self.nasty:
                rr[i, (key + np.arange(0, nout, 53)) % nout] = 0.0       # no coverage
                rr[i, (key + np.arange(7, nout, 211)) % nout] = np.nan
                rr[i, (key + np.arange(3, nout, 401)) % nout] = -0.5     # log2 -> NaN
                rr[i, (key + np.arange(5, nout, 997)) % nout] = np.inf
                rr[i, (key + np.arange(11, nout, 89)) % nout] = 1.0      # log ratio exactly 0 before the median shift
                nn[i, (key + np.arange(1, nout, 61)) % nout] = 20.0      # too few reference bins
                a = (key * 31) % max(1, nout - 400)
                rr[i, a:a + 300] = 0.0                                   # a long run without data
            ml[i] = 0.01 * (key % 7)
            mz[i] = 0.1 * (key % 5)
        self.t += time.perf_counter() - t0
        return 0

    def wcx_predict_stage_ms(self, h, out):
        return 0

    def wcx_cbs_set_boundary(self, h, table, n):
        return 0

    def wcx_cbs_segment(self, h, y, w, off, ns, ids, alpha, nperm, seed, ends, nseg):
        t0 = time.perf_counter()
        offs = _arr(off, (ns + 1,), np.int64)
        total = int(offs[-1])
        yy = _arr(y, (max(total, 1),), np.float64)
        cid = _arr(ids, (max(ns, 1),), np.int32)
        e = _arr(ends, (max(total, 1),), np.int32)
        c = _arr(nseg, (max(ns, 1),), np.int32)
        o = 0
        for s in range(ns):
            n = int(offs[s + 1] - offs[s])
            # cuts from the series itself (and its stream id), not from its position in the call
            key = int(abs(yy[offs[s]:offs[s + 1]][::13]).sum() * 1e6) + int(cid[s])
            cuts = [n] if key % 4 or n < 30 else sorted({1 + key % (n - 1), 1 + (key // 7) % (n - 1), n})
            e[o:o + len(cuts)] = cuts
            c[s] = len(cuts)
            o += len(cuts)
        self.t += time.perf_counter() - t0
        return 0

    def wcx_cbs_stats(self, h, out):
        return 0

    def wcx_segment_zscore(self, h, nr, rows, m, infl, r, w, nb, se, sr, ns, out):
        if ns:
            rr = _arr(r, (nb,), np.float64)
            ww = _arr(w, (nb,), np.float64)
            ii = _arr(infl, (nb,), np.int32)
            seg = _arr(se, (ns, 2), np.int64)
            o = _arr(out, (ns,), np.float64)
            for j in range(ns):  # something that depends on exactly the bins of the segment
                a, b = int(seg[j, 0]), int(seg[j, 1])
                o[j] = float(np.sum(rr[a:b] * ww[a:b])) + float(np.sum(ii[a:b] >= 0))
        return 0


def make_ref_file(binsize=15000, k=300, m=100, seed=5):
    """A reference dict with the keys `predict` reads (reference main.py:168-230), ~3 % of the bins masked.  The index /
    distance / PCA arrays are zeros: the stand-in never reads them."""
    rng = np.random.default_rng(seed)
    ref = {"binsize": binsize, "is_nipt": False, "has_male": True, "has_female": True, "trained_cutoff": 0.006}
    n_aut = None
    base = rng.random(int(synth.bins_per_chr(binsize, 24).sum())) > 0.03
    for g, nchr in (("", 22), (".F", 23), (".M", 24)):
        bpc = synth.bins_per_chr(binsize, nchr)
        mask = base[:int(bpc.sum())].copy()
        if g:  # the gonosomal passes of newref drop a few more autosomal bins AFTER the autosomal snapshot
            mask[rng.choice(np.flatnonzero(mask[:int(bpc[:22].sum())]), 3 if g == ".M" else 5, replace=False)] = False
        offs = np.concatenate([[0], np.cumsum(bpc)])
        per = np.array([int(mask[offs[c]:offs[c + 1]].sum()) for c in range(nchr)], dtype=np.int64)
        n = int(per.sum())
        ref["bins_per_chr" + g] = bpc
        ref["mask" + g] = mask
        ref["masked_bins_per_chr" + g] = per
        ref["masked_bins_per_chr_cum" + g] = np.cumsum(per)
        ref["indexes" + g] = np.zeros((n, k), dtype=np.int32)
        ref["distances" + g] = np.zeros((n, k), dtype=np.float64)
        ref["pca_components" + g] = np.zeros((5, n))
        ref["pca_mean" + g] = np.zeros(n)
        if g == "":
            n_aut = n
        ref["null_ratios" + g] = rng.standard_normal((n, m)) * 0.05
    return ref, n_aut


