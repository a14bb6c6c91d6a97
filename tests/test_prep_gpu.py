"""GPU parity tests of the newref preparation chain (normalize_and_mask, train_pca, PCA-distance filter with the
in-place mask edit and the redo): the kernels against the NumPy oracle (which tests/test_oracle_pin.py pins to the
live reference) and `newref_control.tool_newref_prep` against the golden written by the live reference's own
tool_newref_prep (tests/golden/prep.npz, A -> F -> M with the leaking mask)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import np_oracle  # noqa: E402
from wisecondorx_b200 import newref_control, newref_tools, synth  # noqa: E402


def test_normalize_and_mask_bit_exact():
    samples, _ = synth.make_samples(12, 2_000_000, seed=3, depth=2e6)
    chrs = range(1, 23)
    total = sum(len(samples[0][str(c)]) for c in chrs)
    rng = np.random.default_rng(1)
    mask = rng.random(total) > 0.1
    got = newref_tools.normalize_and_mask(samples, chrs, mask)
    want = np_oracle.normalize_and_mask(samples, chrs, mask)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("n,s", [(3000, 24), (20000, 100), (5000, 129)])
def test_train_pca_matches_exact_pca(n, s):
    rng = np.random.default_rng(7)
    # 7 separated factors around a per-bin profile
    prof = rng.gamma(20.0, 1 / 20.0, n) / n
    f = rng.standard_normal((7, n))
    amps = 0.10 * 0.6 ** np.arange(7)
    x = np.stack([prof * (1 + (amps * rng.standard_normal(7)) @ f) * (1 + 0.01 * rng.standard_normal(n)) for _ in range(s)], axis=1)
    corrected, pca = newref_tools.train_pca(x)
    want_c, comps, mean = np_oracle.train_pca(x)
    np.testing.assert_allclose(pca.mean_, mean, rtol=1e-12)
    np.testing.assert_allclose(pca.components_, comps, rtol=0, atol=1e-8)
    np.testing.assert_allclose(corrected, want_c, rtol=1e-9)
    d, med = newref_tools.pca_distance(corrected)
    bad, cutoff, want_d = np_oracle.pca_distance_filter(want_c)
    np.testing.assert_allclose(med, np.median(corrected, axis=0), rtol=0, atol=0)
    np.testing.assert_allclose(d, want_d, rtol=1e-7)


def test_device_resident_prep_equals_host_functions():
    """newref_tools.DevicePrep (matrices kept in HBM between the steps) returns what the three drop-in functions
    return through host memory: same kernels, same bits."""
    samples, _ = synth.make_samples(14, 1_000_000, seed=9, depth=3e6)
    chrs = range(1, 23)
    counts = newref_tools.stack_counts(samples, chrs)
    rng = np.random.default_rng(2)
    mask = (rng.random(counts.shape[0]) > 0.05) & (counts.sum(axis=1) > 0)  # dead bins would give 0 / 0 rows
    masked = newref_tools.normalize_and_mask(samples, chrs, mask)
    corrected, pca = newref_tools.train_pca(masked)
    d, med = newref_tools.pca_distance(corrected)
    dp = newref_tools.DevicePrep(0)
    assert tuple(dp.normalize_and_mask(counts, mask)) == masked.shape
    assert np.array_equal(dp.fetch("masked"), masked)
    pca2 = dp.train_pca()
    assert np.array_equal(pca2.components_, pca.components_) and np.array_equal(pca2.mean_, pca.mean_)
    assert np.array_equal(dp.fetch("corrected"), corrected, equal_nan=True)  # dead bins give 0 / 0
    d2, med2 = dp.pca_distance()
    assert np.array_equal(d2, d, equal_nan=True) and np.array_equal(med2, med, equal_nan=True)
    # get_reference from the resident corrected matrix == from the host copy
    per = [int(x) for x in np.diff(np.concatenate([[0], np.cumsum([len(samples[0][str(c)]) for c in chrs])]))]
    offs = np.concatenate([[0], np.cumsum(per)])
    mper = np.array([int(mask[offs[i]:offs[i + 1]].sum()) for i in range(22)])
    good = ~np.isnan(corrected).any(axis=1)
    if good.all():
        eng = newref_tools.NewrefEngine(0)
        dp.load_into(eng, mper, np.cumsum(mper))
        i1, d1 = eng.topk(0, masked.shape[0], 50)
        eng.load(corrected, mper, np.cumsum(mper))
        i2, d2 = eng.topk(0, masked.shape[0], 50)
        assert np.array_equal(i1, i2) and np.array_equal(d1, d2)


def test_tool_newref_prep_matches_reference_golden(golden_dir):
    """a1-a3 through the product path: masks / bin counts exact, PCA model and corrected matrix within 1e-5 of the
    reference's sklearn fit (np.random.seed pinned when the golden was made)."""
    g = np.load(os.path.join(golden_dir, "prep.npz"))
    offs = np.concatenate([[0], np.cumsum(g["lens"])])
    genders = [str(x) for x in g["genders"]]
    samples = []
    for i in range(g["counts"].shape[1]):
        s = {str(c + 1): g["counts"][offs[c]:offs[c + 1], i].copy() for c in range(24)}
        if genders[i] == "M":  # gender_correct, main.py:95-97
            s["23"], s["24"] = s["23"] * 2, s["24"] * 2
        samples.append(s)
    bpc = [int(x) for x in g["bins_per_chr"]]
    total_mask = g["total_mask"].copy()
    for gender in ("A", "F", "M"):
        sub = [s for s, gg in zip(samples, genders) if gender == "A" or gg == gender]
        prep = newref_control.tool_newref_prep(sub, gender, total_mask, bpc)
        assert np.array_equal(prep["mask"], g[gender + "_mask"])
        assert np.array_equal(total_mask, g[gender + "_total_mask_after"])  # the edit leaked into the caller's mask
        assert np.array_equal(prep["masked_bins_per_chr"], g[gender + "_masked_bins_per_chr"])
        assert np.array_equal(prep["masked_bins_per_chr_cum"], g[gender + "_masked_bins_per_chr_cum"])
        np.testing.assert_allclose(prep["pca_mean"], g[gender + "_pca_mean"], rtol=1e-12)
        np.testing.assert_allclose(prep["pca_components"], g[gender + "_pca_components"], rtol=0, atol=1e-5)
        corrected = prep["pca_corrected_data"].fetch("corrected")
        np.testing.assert_allclose(corrected[::5], g[gender + "_corrected_rows"], rtol=1e-5)


def test_train_pca_rank_deficient_pass():
    """A gonosomal pass may run with exactly 5 samples (main.py:104,119): the centred matrix has rank 4, so the
    fifth component must not be divided by a round-off singular value (it used to reach ~1e130 and turn every
    predicted ratio into ~1e-257).  Rows are unit norm and mutually orthogonal, the correction equals the rank-4
    exact PCA and a projection of a new sample stays finite."""
    rng = np.random.default_rng(5)
    n, s = 4000, 5
    prof = rng.gamma(20.0, 1 / 20.0, n) / n
    x = np.stack([prof * (1 + 0.05 * rng.standard_normal(n)) for _ in range(s)], axis=1)
    corrected, pca = newref_tools.train_pca(x)
    assert pca.components_.shape == (5, n) and np.isfinite(pca.components_).all()
    gram = pca.components_ @ pca.components_.T
    np.testing.assert_allclose(gram, np.eye(5), atol=1e-9)
    want_c, comps4, mean = np_oracle.train_pca(x, 4)
    np.testing.assert_allclose(corrected, want_c, rtol=1e-9)
    np.testing.assert_allclose(pca.components_[:4], comps4, rtol=0, atol=1e-8)
    new = prof * (1 + 0.05 * rng.standard_normal(n))
    rec = (new - pca.mean_) @ pca.components_.T @ pca.components_ + pca.mean_
    assert np.isfinite(new / rec).all() and np.abs(new / rec).max() < 10
    dp = newref_tools.DevicePrep(0)
    # the device-resident chain takes the same path
    counts = np.round(x * 4e9).astype(np.int32)
    dp.normalize_and_mask(counts, np.ones(n, dtype=bool))
    pca2 = dp.train_pca()
    assert np.isfinite(pca2.components_).all() and np.allclose(np.linalg.norm(pca2.components_, axis=1), 1.0)
