"""Parity AT THE BENCH CONFIGURATION (BASELINE config 3: 500 samples @ 15 kb, 191 678 autosomal bins, refsize 300) and
at the predict configuration built on it (config 4: one sample against that reference).  The CPU oracle cannot run all
3.5e10 pairs, so the GPU result of the WHOLE pass is compared on row windows spread over the genome (every candidate
of those rows is evaluated by the oracle): indexes and distances bit-exact, null ratios 1e-12.  This is where the
4-leaf summation plan of S = 500, WCX_CAND_CAP, int32 slot arithmetic over 12.6 GB of candidate lists and the 8-block
side-stream path are exercised.  `normalize` at this size is checked in full against the NumPy oracle."""
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import c_oracle, np_oracle  # noqa: E402
from wisecondorx_b200 import _lib, newref_tools, predict_control, predict_tools, synth  # noqa: E402

K, M = 300, 100


@pytest.fixture(scope="module")
def config3():
    per = synth.config_bins(3)
    x, per, cum = synth.make_corrected_matrix(per, 500, seed=3)
    n = x.shape[0]
    eng = newref_tools.NewrefEngine(0, _lib.Context(0))
    ids = np.arange(M, dtype=np.int32)
    idx, dist, nr = eng.get_reference_host(x, per, cum, 0, n, K, ids)
    st = eng.stats()
    yield {"x": x, "per": per, "cum": cum, "idx": idx, "dist": dist, "nr": nr, "ids": ids, "stats": st}
    eng.ctx.close()


def test_config3_row_windows_vs_c_oracle(config3):
    c = config3
    n = c["x"].shape[0]
    assert n == 191678 and c["x"].shape[1] == 500
    rng = np.random.default_rng(33)
    # windows: both ends of the genome, chromosome boundaries (the excluded range changes inside the window), random
    starts = [0, n - 16, int(c["cum"][0]) - 8, int(c["cum"][10]) - 8] + [int(v) for v in rng.integers(0, n - 16, 8)]
    checked = 0
    for s0 in starts:
        e0 = s0 + 16
        oi, od = c_oracle.topk(c["x"], c["per"], c["cum"], K, s0, e0)
        assert np.array_equal(c["idx"][s0:e0], oi), s0
        assert np.array_equal(c["dist"][s0:e0], od), s0
        onr = c_oracle.null_ratios(c["x"], oi, s0, e0, c["ids"])
        np.testing.assert_allclose(c["nr"][s0:e0], onr, rtol=1e-12, atol=1e-14)
        checked += e0 - s0
    assert checked == 16 * len(starts)
    # structural properties over ALL rows: ascending distances, positions inside the chr-excluded range, no self hits
    assert (np.diff(c["dist"], axis=1) >= 0).all()
    chrom = np.searchsorted(c["cum"], np.arange(n), side="right")
    nex = n - c["per"][chrom]
    assert (c["idx"] >= 0).all() and (c["idx"] < nex[:, None]).all()
    assert np.isfinite(c["nr"]).all()
    assert c["stats"]["exact_fallback_rows"] < n // 1000, c["stats"]


def test_config3_parts_equal_whole(config3):
    """A part of the reference's fan-out (newref_tools.py:244-247) returns the rows of the whole pass (multi-GPU shards
    are exactly these parts)."""
    c = config3
    n = c["x"].shape[0]
    eng = newref_tools.NewrefEngine(0, _lib.Context(0))
    eng.load(c["x"], c["per"], c["cum"])
    for part, parts in ((3, 8), (8, 8)):
        s0, e0 = newref_tools._get_part(part - 1, parts, n)
        idx, dist, nr = eng.reference(s0, e0, K, c["ids"])
        assert np.array_equal(idx, c["idx"][s0:e0]) and np.array_equal(dist, c["dist"][s0:e0])
        assert np.array_equal(nr, c["nr"][s0:e0])
    eng.ctx.close()


def test_config4_normalize_15kb_vs_oracle(config3):
    """BASELINE config 4: `normalize` (coverage + PCA projection + three within-sample passes + the two medians) of
    one sample against the config-3 reference, every bin compared with np_oracle.normalize (pinned to the live
    reference by tests/test_oracle_pin.py / test_oracle_golden.py)."""
    c = config3
    n = c["x"].shape[0]
    rng = np.random.default_rng(44)
    comps = np.linalg.qr(rng.standard_normal((n, 5)))[0].T.copy()
    extra = rng.multinomial(300, c["per"] / n)  # 300 masked-out bins spread over the chromosomes: the gather through mask_pos
    bpc = c["per"] + extra
    mask = np.ones(int(bpc.sum()), dtype=bool)
    o = 0
    for nb, ex in zip(bpc, extra):
        mask[o + rng.choice(int(nb), int(ex), replace=False)] = False
        o += int(nb)
    assert int(mask.sum()) == n
    ref = {"indexes": c["idx"], "distances": c["dist"], "masked_bins_per_chr": c["per"], "masked_bins_per_chr_cum": c["cum"],
           "pca_components": comps, "pca_mean": np.full(n, 1.0 / n), "mask": mask, "bins_per_chr": bpc}
    offs = np.concatenate([[0], np.cumsum(bpc)]).astype(int)
    lam = np.zeros(n + 300)
    lam[mask] = 60.0 * np.clip(c["x"][:, 7], 0, None)
    lam[offs[4] + 2000: offs[4] + 2400] *= 1.5
    counts = rng.poisson(lam).astype(np.int32)
    sample = {str(k + 1): counts[offs[k]:offs[k + 1]] for k in range(22)}
    args = types.SimpleNamespace(maskrepeats=5)
    eng = predict_tools.PredictEngine(0, _lib.Context(0))
    r, z, w, nref, m_lr, m_z = predict_control.normalize(args, sample, ref, "A", eng)
    wr, wz, ww, wn, wm_lr, wm_z = np_oracle.normalize(sample, ref, "A", 5)
    assert np.array_equal(nref, wn)
    np.testing.assert_allclose(w, ww, rtol=1e-12)
    np.testing.assert_allclose(r, wr, rtol=1e-9, equal_nan=True)
    np.testing.assert_allclose(z, wz, rtol=1e-6, atol=1e-8, equal_nan=True)
    np.testing.assert_allclose([m_lr, m_z], [wm_lr, wm_z], rtol=1e-9, atol=1e-12)
    eng.ctx.close()
