"""Deterministic synthetic inputs for tests and bench.

Two generators:

* ``make_samples``  -- per-sample read-count dicts with the sample ``.npz`` schema the
  reference's ``convert`` writes (reference ``main.py:33-35``): keys "1".."24" -> int32[bins].
* ``make_corrected_matrix`` -- a [N, S] float64 matrix shaped like the reference's
  ``pca_corrected_data`` (``newref_control.py:68``), i.e. ratios around 1.0 with bin-specific
  noise plus masked-bin counts per chromosome.  This is the parity/bench injection point for
  ``get_reference`` ("checkpoint/resume" row).

No file under ``/root/reference`` is read here; this module travels to the GPU box.
"""
from __future__ import annotations

import numpy as np

# hg38 primary-assembly chromosome lengths, chr1..22, X, Y
HG38_LENGTHS = [
    248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973,
    145138636, 138394717, 133797422, 135086622, 133275309, 114364328, 107043718,
    101991189, 90338345, 83257441, 80373285, 58617616, 64444167, 46709983, 50818468,
    156040895, 57227415,
]


def bins_per_chr(binsize: float, nchr: int = 24) -> np.ndarray:
    """Bins per chromosome as the reference's convert step lays them out
    (``convert_tools.py:63-65``: ``int(len / binsize) + 1``)."""
    return np.array([int(l / binsize) + 1 for l in HG38_LENGTHS[:nchr]], dtype=np.int64)


def make_samples(n_samples: int, binsize: float, seed: int, depth: float = 6e6,
                 n_factors: int = 7, dead_frac: float = 0.03, cnv=None):
    """Poisson read counts with ``n_factors`` latent bias factors of geometrically decaying
    amplitude (0.10 * 0.6**f) so the top-5 principal components are identifiable.

    Returns (samples, genders): list of dicts "1".."24" -> int32 array, list of "F"/"M".
    ``cnv``: optional list of (sample_idx, chr(1-based), start_bin, end_bin, ratio).
    """
    rng = np.random.default_rng(seed)
    bpc = bins_per_chr(binsize)
    total = int(bpc.sum())
    profile = rng.gamma(shape=20.0, scale=1.0 / 20.0, size=total)
    dead = rng.random(total) < dead_frac
    profile[dead] = 0.0
    factors = rng.standard_normal((n_factors, total))
    amps = 0.10 * 0.6 ** np.arange(n_factors)
    genders = ["F" if i % 2 == 0 else "M" for i in range(n_samples)]
    offs = np.concatenate([[0], np.cumsum(bpc)])
    samples = []
    for i in range(n_samples):
        g = rng.standard_normal(n_factors)
        lam = profile * (1.0 + (amps * g) @ factors)
        lam = np.clip(lam, 0.0, None)
        cn = np.ones(total)
        if genders[i] == "M":
            cn[offs[22]:offs[23]] = 0.5
            cn[offs[23]:offs[24]] = 0.5
        else:
            # females keep a little chrY coverage (X-homolog mis-mapping), see SURVEY A.4
            cn[offs[23]:offs[24]] = 0.2
        if cnv:
            for (si, c, s, e, ratio) in cnv:
                if si == i:
                    cn[offs[c - 1] + s: offs[c - 1] + e] *= ratio
        lam = lam * cn
        lam = lam / lam.sum() * depth * (0.8 + 0.4 * rng.random())
        counts = rng.poisson(lam).astype(np.int32)
        samples.append({str(c + 1): counts[offs[c]:offs[c + 1]].copy() for c in range(24)})
    return samples, genders


def make_corrected_matrix(n_chr_bins, n_samples: int, seed: int, noise: float = 0.05,
                          n_factors: int = 3, dtype=np.float64):
    """A ``pca_corrected_data``-like matrix: 1 + residual structure + per-bin noise.

    ``n_chr_bins``: masked bins per chromosome (list).  Returns (X[N,S] C-order float64,
    masked_bins_per_chr int64, masked_bins_per_chr_cum int64).
    Residual low-rank structure gives bins genuinely similar neighbours so that the top-k
    selection is not a pure noise ranking.
    """
    rng = np.random.default_rng(seed)
    per = np.asarray(n_chr_bins, dtype=np.int64)
    n = int(per.sum())
    # process in chunks to bound temporary memory at large N*S
    x = np.empty((n, n_samples), dtype=np.float64)
    load = rng.standard_normal((n_factors, n_samples))
    bin_sigma = noise * (0.6 + 0.8 * rng.random(n))
    chunk = 1 << 15
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        w = 0.02 * rng.standard_normal((e - s, n_factors))
        x[s:e] = 1.0 + w @ load + bin_sigma[s:e, None] * rng.standard_normal((e - s, n_samples))
    return x.astype(dtype, copy=False), per, np.cumsum(per)


def config_bins(config: int):
    """Nominal (unmasked) autosomal bins per chromosome for the BASELINE.json configs:
    1 -> 1 Mb, 2 -> 100 kb, 3 -> 15 kb."""
    binsize = {1: 1e6, 2: 1e5, 3: 15e3}[config]
    return bins_per_chr(binsize, 22)
