"""GPU parity tests for the predict path (normalize, get_weights, get_optimal_cutoff, get_z_score)
through the C-ABI against the golden vectors of the live reference and the NumPy oracle.
Tolerance: 1e-5 (BASELINE.json north_star) -- asserted tighter where the arithmetic allows."""
import os
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import np_oracle  # noqa: E402
from wisecondorx_b200 import predict_control, predict_tools  # noqa: E402


@pytest.fixture(scope="module")
def gpred(golden_dir):
    return np.load(os.path.join(golden_dir, "newref_predict.npz"), allow_pickle=True)


@pytest.fixture(scope="module")
def ref(gpred):
    return {k[5:]: gpred[k] for k in gpred.files if k.startswith("ref__")}


def _sample(gpred, si, g):
    sample = {str(c): gpred[f"t{si}_sample_{c}"].copy() for c in range(1, 25)}
    if g == "M":  # gender_correct, overall_tools.py:48-53
        sample["23"] = sample["23"] * 2
        sample["24"] = sample["24"] * 2
    return sample


def test_weights_and_cutoff(gpred, ref):
    eng = predict_tools.PredictEngine(0)
    for ap in ["", ".F", ".M"]:
        w = eng.get_weights(ref, ap)
        np.testing.assert_allclose(w, np_oracle.get_weights(ref["distances" + ap]), rtol=1e-12)
    c = eng.get_optimal_cutoff(ref, 5)
    np.testing.assert_allclose(c, gpred["t0_cutoff"], rtol=1e-12)


@pytest.mark.parametrize("si,g", [(0, "F"), (1, "M")])
def test_normalize_matches_reference_golden(gpred, ref, si, g):
    args = types.SimpleNamespace(maskrepeats=5)
    sample = _sample(gpred, si, g)
    eng = predict_tools.PredictEngine(0)
    for rg in ["A", g]:
        r, z, w, n, m_lr, m_z = predict_control.normalize(args, sample, ref, rg, engine=eng)
        np.testing.assert_allclose(r, gpred[f"t{si}_{rg}_r"], rtol=1e-9, equal_nan=True)
        np.testing.assert_allclose(z, gpred[f"t{si}_{rg}_z"], rtol=1e-6, atol=1e-8, equal_nan=True)
        np.testing.assert_allclose(w, gpred[f"t{si}_{rg}_w"], rtol=1e-12)
        assert np.array_equal(n, gpred[f"t{si}_{rg}_n"])
        np.testing.assert_allclose([m_lr, m_z], gpred[f"t{si}_{rg}_m"], rtol=1e-7, atol=1e-9)


def test_normalize_batch_equals_single(gpred, ref):
    args = types.SimpleNamespace(maskrepeats=5)
    eng = predict_tools.PredictEngine(0)
    s0, s1 = _sample(gpred, 0, "F"), _sample(gpred, 1, "F")
    rb, zb, wb, nb, mlb, mzb = predict_control.normalize_batch(args, [s0, s1, s0], ref, "A", engine=eng)
    for i, s in enumerate([s0, s1, s0]):
        r, z, w, n, m_lr, m_z = predict_control.normalize(args, s, ref, "A", engine=eng)
        assert np.array_equal(rb[i], r, equal_nan=True) and np.array_equal(zb[i], z, equal_nan=True)
        assert np.array_equal(nb[i], n) and mlb[i] == m_lr and mzb[i] == m_z


def test_get_z_score_matches_reference_golden(gpred):
    bpc = gpred["zs_bpc"]
    offs = np.concatenate([[0], np.cumsum(bpc)])
    r, w, nr, has = gpred["zs_r"], gpred["zs_w"], gpred["zs_nr"], gpred["zs_has_nr"]
    res = {"results_r": [r[offs[c]:offs[c + 1]].tolist() for c in range(len(bpc))],
           "results_w": [w[offs[c]:offs[c + 1]].tolist() for c in range(len(bpc))],
           "results_nr": [[nr[i].tolist() if has[i] else 0 for i in range(offs[c], offs[c + 1])] for c in range(len(bpc))]}
    segs = [[int(s[0]), int(s[1]), int(s[2]), float(s[3])] for s in gpred["zs_segs"]]
    zs = predict_tools.get_z_score(segs, res)
    got = np.array([np.nan if isinstance(z, str) else z for z in zs])
    np.testing.assert_allclose(got, gpred["zs_z"], rtol=1e-9, equal_nan=True)


def test_normalize_midsize_vs_oracle():
    """100 kb-like layout: reference built by the oracle, normalisation compared bin by bin."""
    from wisecondorx_b200 import synth
    rng = np.random.default_rng(5)
    per = (synth.config_bins(2) // 8).astype(np.int64)
    per = np.concatenate([per, [200, 70]])
    n = int(per.sum()); cum = np.cumsum(per); k = 120
    # synthetic reference arrays with the right structure
    idx = np.empty((n, k), dtype=np.int32); dist = np.empty((n, k))
    chrom = np.searchsorted(cum, np.arange(n), side="right")
    for i in range(n):
        nex = n - per[chrom[i]]
        idx[i] = rng.choice(nex, k, replace=False)
        dist[i] = np.sort(rng.gamma(4.0, 0.02, k))
    idx[7, -3:] = -1; dist[7, -3:] = 1e10
    comps = np.linalg.qr(rng.standard_normal((n, 5)))[0].T.copy()
    mean = np.abs(rng.normal(1.0 / n, 0.1 / n, n))
    mask = np.ones(n + 300, dtype=bool); mask[rng.choice(n + 300, 300, replace=False)] = False
    bins_total = len(mask)
    bpc = np.diff(np.concatenate([[0], np.round(np.linspace(0, bins_total, len(per) + 1)[1:]).astype(int)]))
    ref = {"indexes": idx, "distances": dist, "masked_bins_per_chr": per, "masked_bins_per_chr_cum": cum,
           "pca_components": comps, "pca_mean": mean, "mask": mask, "bins_per_chr": bpc}
    for sfx in (".M",):
        for key in list(ref):
            ref[key + sfx] = ref[key]
    samples = []
    for b in range(3):
        counts = rng.poisson(60, bins_total).astype(np.int32)
        counts[rng.choice(bins_total, 50)] = 0
        offs = np.concatenate([[0], np.cumsum(bpc)])
        samples.append({str(c + 1): counts[offs[c]:offs[c + 1]] for c in range(len(bpc))})
    args = types.SimpleNamespace(maskrepeats=5)
    eng = predict_tools.PredictEngine(0)
    for rg in ("A", "M"):
        rb, zb, wb, nb, mlb, mzb = predict_control.normalize_batch(args, samples, ref, rg, engine=eng)
        for b, s in enumerate(samples):
            r, z, w, nn, m_lr, m_z = np_oracle.normalize(s, ref, rg)
            np.testing.assert_allclose(rb[b], r, rtol=1e-9, equal_nan=True)
            np.testing.assert_allclose(zb[b], z, rtol=1e-6, atol=1e-8, equal_nan=True)
            assert np.array_equal(nb[b], nn)
            np.testing.assert_allclose([mlb[b], mzb[b]], [m_lr, m_z], rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(wb, w, rtol=1e-12)
