"""GPU parity tests of the newref preparation kernels (normalize_and_mask, train_pca, PCA-distance
filter) against the NumPy oracle; train_pca additionally against the live reference's golden
where the problem is well conditioned."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import np_oracle  # noqa: E402
from wisecondorx_b200 import newref_tools, synth  # noqa: E402


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
