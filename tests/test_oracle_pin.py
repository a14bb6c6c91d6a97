"""Pins oracle/np_oracle.py (and the C restatement) against the LIVE reference on fresh seeded inputs -- other
seeds and shapes than the committed golden vectors.  Runs only where /root/reference exists (the build
container); on the GPU box the committed fixtures (tests/test_oracle_golden.py) are the pin."""
import random
import types

import numpy as np
import pytest

from oracle import c_oracle, np_oracle, ref_loader
from wisecondorx_b200 import synth

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="live reference not present (GPU box)")


@pytest.fixture(scope="module")
def R():
    return ref_loader.load()


@pytest.mark.parametrize("seed,s,k,part,parts,gon", [(101, 17, 25, 1, 1, False), (102, 40, 30, 2, 4, False),
                                                      (103, 9, 15, 1, 1, True), (104, 130, 20, 3, 3, False)])
def test_get_reference_live(R, seed, s, k, part, parts, gon):
    per = [31, 27, 24, 22, 20, 18, 16, 15, 14, 13, 12, 11, 10, 9, 9, 8, 8, 7, 6, 6, 5, 5] + ([14, 5] if gon else [])
    x, per, cum = synth.make_corrected_matrix(per, s, seed=seed)
    if seed == 102:
        x = np.round(x * 32) / 32  # exact ties
    random.seed(seed)
    want = R.newref_tools.get_reference(x, per, cum, k, part, parts)
    random.seed(seed)
    ids = random.sample(range(s), min(s, 100))
    got = np_oracle.get_reference(x, per, cum, k, part, parts, ids)
    assert np.array_equal(got[0], want[0]) and got[0].dtype == want[0].dtype
    assert np.array_equal(got[1], want[1])
    np.testing.assert_allclose(got[2], want[2], rtol=1e-13, atol=0, equal_nan=True)
    start, end = np_oracle.get_part(part - 1, parts, int(cum[-1]))
    ci, cd = c_oracle.topk(x, per, cum, k, start, end)
    assert np.array_equal(ci, want[0]) and np.array_equal(cd, want[1])
    np.testing.assert_allclose(c_oracle.null_ratios(x, ci, start, end, ids), want[2], rtol=1e-13, atol=0, equal_nan=True)


def test_predict_pieces_live(R):
    rng = np.random.default_rng(7)
    n, k = 400, 30
    dist = np.sort(rng.random((n, k)) * 3.0, axis=1)
    ref_file = {"distances": dist}
    np.testing.assert_allclose(np_oracle.get_weights(dist), R.predict_tools.get_weights(ref_file, ""), rtol=1e-14)
    for rep in (1, 5):
        assert np.isclose(np_oracle.get_optimal_cutoff(dist, rep), R.predict_tools.get_optimal_cutoff(ref_file, rep), rtol=1e-14)
    # normalize_repeat on a small reference: per-chromosome layout, chr-excluded indexes
    per = np.array([60, 50, 40, 30] + [10] * 18 + [12, 8])
    cum = np.cumsum(per)
    ntot = int(cum[-1])
    idx = np.empty((ntot, k), dtype=np.int32)
    for c in range(len(per)):
        s0, e0 = int(cum[c] - per[c]), int(cum[c])
        idx[s0:e0] = rng.integers(0, ntot - per[c], size=(e0 - s0, k))
    d2 = np.sort(rng.random((ntot, k)), axis=1)
    test = np.abs(1.0 + 0.1 * rng.standard_normal(ntot))
    test[5] = 3.0  # an outlier that gets masked after the first pass
    ref = {"indexes": idx, "distances": d2, "masked_bins_per_chr": per, "masked_bins_per_chr_cum": cum}
    for ap, ct, cp in (("", 0, 0), (".F", int(cum[21]), 22)):
        rf = {"indexes" + ap: idx, "distances" + ap: d2, "masked_bins_per_chr" + ap: per, "masked_bins_per_chr_cum" + ap: cum}
        want = R.predict_tools.normalize_repeat(test.copy(), rf, 0.8, ct, cp, ap)
        got = np_oracle.normalize_repeat(test.copy(), idx, d2, per, cum, 0.8, ct, cp)
        np.testing.assert_allclose(got[0], want[0], rtol=1e-10, equal_nan=True)   # z
        np.testing.assert_allclose(got[1], want[1], rtol=1e-12, equal_nan=True)   # r
        assert np.array_equal(got[2], want[2])                                    # ref sizes
        assert np.isclose(got[3], want[3], rtol=1e-12) and np.isclose(got[4], want[4], rtol=1e-10)


def _planted_samples(S, binsize, seed):
    samples, genders = synth.make_samples(S, binsize, seed=seed, depth=8e6)
    rng = np.random.default_rng(seed + 1)
    offs = np.concatenate([[0], np.cumsum([len(samples[0][str(c)]) for c in range(1, 25)])])
    for b in rng.choice(int(offs[22]), 6, replace=False):  # bins the five components cannot explain
        c = int(np.searchsorted(offs, b, side="right")) - 1
        for s in samples:
            s[str(c + 1)][b - offs[c]] = int(s[str(c + 1)][b - offs[c]] * rng.lognormal(0, 1.2))
    return samples, genders


@pytest.mark.parametrize("S,binsize,seed", [(100, 1_000_000, 41), (24, 5_000_000, 7)])
def test_prep_chain_live(R, tmp_path, S, binsize, seed):
    """a1-a3: normalize_and_mask (bit-exact), get_mask, train_pca (np.random.seed pinned, 1e-5 per SURVEY A.3) and
    the PCA-distance filter INCLUDING the in-place mask edit and the redo (newref_control.py:38-58) against the
    live reference, A -> F -> M with the mask leaking from pass to pass (main.py:98-137)."""
    samples, genders = _planted_samples(S, binsize, seed)
    samples = np.array(samples)
    for i, s in enumerate(samples):
        samples[i] = R.overall_tools.gender_correct(s, genders[i])
    mask_ref, bpc = R.newref_tools.get_mask(samples)
    mask_o, bpc_o = np_oracle.get_mask(samples)
    assert np.array_equal(mask_ref, mask_o) and list(bpc) == list(bpc_o)
    want_nm = R.newref_tools.normalize_and_mask(samples, range(1, 23), mask_ref[: sum(bpc[:22])])
    assert np.array_equal(np_oracle.normalize_and_mask(samples, range(1, 23), mask_ref[: sum(bpc[:22])]), want_nm)
    m1, m2 = mask_ref.copy(), mask_ref.copy()
    g = np.array(genders)
    removed = 0
    for gender, sub in (("A", samples), ("F", samples[g == "F"]), ("M", samples[g == "M"])):
        a = types.SimpleNamespace(prepdatafile=str(tmp_path / "d.npy"), prepfile=str(tmp_path / "p.npz"), binsize=binsize)
        np.random.seed(3)
        R.newref_control.tool_newref_prep(a, sub, gender, m1, bpc)
        want, wc = np.load(a.prepfile), np.load(a.prepdatafile)
        got = np_oracle.tool_newref_prep(sub, gender, m2, bpc)
        removed += got["n_removed"]
        assert np.array_equal(m1, m2)  # the leak into the caller's mask
        assert np.array_equal(want["mask"], got["mask"])
        assert np.array_equal(want["masked_bins_per_chr"], got["masked_bins_per_chr"])
        assert np.array_equal(want["masked_bins_per_chr_cum"], got["masked_bins_per_chr_cum"])
        np.testing.assert_allclose(got["pca_mean"], want["pca_mean"], rtol=1e-12)
        np.testing.assert_allclose(got["pca_components"], want["pca_components"], rtol=0, atol=1e-5)
        np.testing.assert_allclose(got["pca_corrected_data"], wc, rtol=1e-5)
    assert removed > 0  # the filter and the redo were exercised


def test_qc_per_bin_stats_live(R):
    rng = np.random.default_rng(3)
    idx = rng.integers(0, 1000, (500, 40)).astype(np.int32)
    dist = np.sort(rng.random((500, 40)) * 4, axis=1)
    want = R.ref_qc._compute_per_bin_stats(idx, dist)
    got = np_oracle.per_bin_stats(idx, dist)
    for a, b in zip(got, want):
        np.testing.assert_allclose(a, b, rtol=1e-15)
