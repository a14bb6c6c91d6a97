"""CPU suite: the NumPy oracle (oracle/np_oracle.py) against the golden vectors produced by the
live reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import np_oracle as O


@pytest.fixture(scope="module")
def gref(golden_dir):
    return np.load(os.path.join(golden_dir, "get_reference.npz"))


@pytest.fixture(scope="module")
def gpred(golden_dir):
    return np.load(os.path.join(golden_dir, "newref_predict.npz"), allow_pickle=True)


@pytest.mark.parametrize("case,part,parts,k", [("A_p11", 1, 1, 30), ("A_p23", 2, 3, 30),
                                                ("G", 1, 1, 30), ("T", 1, 1, 12), ("S", 1, 1, 20)])
def test_get_reference_bit_exact(gref, case, part, parts, k):
    base = case.split("_")[0]
    x, per, cum = gref[base + "_x"], gref[base + "_per"], gref[base + "_cum"]
    idx, dist, nr = O.get_reference(x, per, cum, k, part, parts, gref[case + "_ids"].tolist())
    assert idx.dtype == np.int32
    assert np.array_equal(idx, gref[case + "_idx"])
    assert np.array_equal(dist, gref[case + "_dist"])  # same NumPy expression -> bit exact
    assert np.array_equal(nr, gref[case + "_nr"], equal_nan=True)


def _ref(gpred):
    return {k[5:]: gpred[k] for k in gpred.files if k.startswith("ref__")}


@pytest.mark.parametrize("si,g", [(0, "F"), (1, "M")])
def test_normalize_matches_reference(gpred, si, g):
    ref = _ref(gpred)
    sample = {str(c): gpred[f"t{si}_sample_{c}"].copy() for c in range(1, 25)}
    if g == "M":  # overall_tools.py:48-53
        sample["23"] = sample["23"] * 2
        sample["24"] = sample["24"] * 2
    assert np.isclose(O.get_optimal_cutoff(ref["distances"], 5), gpred[f"t{si}_cutoff"], rtol=1e-12)
    for rg in ["A", g]:
        r, z, w, n, m_lr, m_z = O.normalize(sample, ref, rg)
        np.testing.assert_allclose(r, gpred[f"t{si}_{rg}_r"], rtol=1e-9, equal_nan=True)
        np.testing.assert_allclose(z, gpred[f"t{si}_{rg}_z"], rtol=1e-7, atol=1e-9, equal_nan=True)
        np.testing.assert_allclose(w, gpred[f"t{si}_{rg}_w"], rtol=1e-12)
        assert np.array_equal(n, gpred[f"t{si}_{rg}_n"])
        np.testing.assert_allclose([m_lr, m_z], gpred[f"t{si}_{rg}_m"], rtol=1e-7, atol=1e-10)


def test_get_z_score_matches_reference(gpred):
    bpc = gpred["zs_bpc"]
    offs = np.concatenate([[0], np.cumsum(bpc)])
    r, w, nr, has = gpred["zs_r"], gpred["zs_w"], gpred["zs_nr"], gpred["zs_has_nr"]
    res_r = [r[offs[c]:offs[c + 1]] for c in range(len(bpc))]
    res_w = [w[offs[c]:offs[c + 1]] for c in range(len(bpc))]
    res_nr = [[nr[i] if has[i] else 0 for i in range(offs[c], offs[c + 1])] for c in range(len(bpc))]
    segs = [[int(s[0]), int(s[1]), int(s[2]), float(s[3])] for s in gpred["zs_segs"]]
    zs = O.get_z_score(segs, res_nr, res_r, res_w)
    got = np.array([np.nan if isinstance(z, str) else z for z in zs])
    np.testing.assert_allclose(got, gpred["zs_z"], rtol=1e-9, equal_nan=True)


@pytest.mark.parametrize("case,part,parts,k", [("A_p11", 1, 1, 30), ("A_p23", 2, 3, 30),
                                                ("G", 1, 1, 30), ("T", 1, 1, 12), ("S", 1, 1, 20)])
def test_c_oracle_get_reference(gref, case, part, parts, k):
    from oracle import c_oracle
    c_oracle.build()
    base = case.split("_")[0]
    x, per, cum = gref[base + "_x"], gref[base + "_per"], gref[base + "_cum"]
    idx, dist, nr = c_oracle.get_reference(x, per, cum, k, part, parts, gref[case + "_ids"])
    assert np.array_equal(idx, gref[case + "_idx"])
    assert np.array_equal(dist, gref[case + "_dist"])  # NumPy pairwise order replicated
    np.testing.assert_allclose(nr, gref[case + "_nr"], rtol=1e-13, atol=1e-15, equal_nan=True)


# ---- a15 / f2: result assembly and post-processing against the reference's own tool_test ---------------------
@pytest.fixture(scope="module")
def gtool(golden_dir):
    return np.load(os.path.join(golden_dir, "tool_test.npz"), allow_pickle=True)


@pytest.mark.parametrize("si,g", [(0, "F"), (1, "M")])
def test_assembly_matches_reference_tool_test(gpred, gtool, si, g):
    """normalize x2 -> assembly (main.py:242-271) -> get_post_processed_result -> log_trans -> apply_blacklist,
    restated in np_oracle, equals what the live reference's tool_test handed to generate_output_tables."""
    ref = _ref(gpred)
    sample = {str(c): gpred[f"t{si}_sample_{c}"].copy() for c in range(1, 25)}
    if g == "M":
        sample["23"], sample["24"] = sample["23"] * 2, sample["24"] * 2
    assert gtool[f"t{si}_meta"][0] == g
    aut = O.normalize(sample, ref, "A")
    gon = O.normalize(sample, ref, g)
    sfx = "." + g
    res, _ = O.assemble_results(aut, gon, ref["null_ratios"], ref["null_ratios" + sfx][len(ref["null_ratios"]):], 10,
                                ref["mask" + sfx], ref["bins_per_chr" + sfx])
    O.log_trans(res, aut[4])
    if si == 0:
        O.apply_blacklist(res, str(gtool["blacklist_text"]), int(gtool[f"t{si}_meta"][3]))
    for key, tol in (("results_r", 1e-9), ("results_z", 1e-6), ("results_w", 1e-12)):
        got = np.concatenate(res[key])
        want = gtool[f"t{si}_{key}"]
        assert np.array_equal(got == 0, want == 0), key
        np.testing.assert_allclose(got, want, rtol=tol, atol=1e-10, err_msg=key)


def test_prep_chain_matches_reference_golden(golden_dir):
    """np_oracle.tool_newref_prep (exact PCA) against the live reference's tool_newref_prep (sklearn PCA, seed
    pinned) for the A -> F -> M passes with the leaking mask: masks and bin counts exact, model within 1e-5
    (SURVEY.md A.3)."""
    g = np.load(os.path.join(golden_dir, "prep.npz"))
    samples, genders, bpc = _prep_samples(g)
    total_mask = g["total_mask"].copy()
    for gender in ("A", "F", "M"):
        sub = [s for s, gg in zip(samples, genders) if gender == "A" or gg == gender]
        got = O.tool_newref_prep(sub, gender, total_mask, bpc)
        assert got["n_removed"] > 0 or gender != "A"
        assert np.array_equal(got["mask"], g[gender + "_mask"])
        assert np.array_equal(total_mask, g[gender + "_total_mask_after"])
        assert np.array_equal(got["masked_bins_per_chr"], g[gender + "_masked_bins_per_chr"])
        np.testing.assert_allclose(got["pca_mean"], g[gender + "_pca_mean"], rtol=1e-12)
        np.testing.assert_allclose(got["pca_components"], g[gender + "_pca_components"], rtol=0, atol=1e-5)
        np.testing.assert_allclose(got["pca_corrected_data"][::5], g[gender + "_corrected_rows"], rtol=1e-5)


def _prep_samples(g):
    """Sample dicts of the prep golden (gender-corrected as main.py:95-97 does before get_mask)."""
    offs = np.concatenate([[0], np.cumsum(g["lens"])])
    genders = [str(x) for x in g["genders"]]
    samples = []
    for i in range(g["counts"].shape[1]):
        s = {str(c + 1): g["counts"][offs[c]:offs[c + 1], i].copy() for c in range(24)}
        if genders[i] == "M":
            s["23"], s["24"] = s["23"] * 2, s["24"] * 2
        samples.append(s)
    return samples, genders, [int(x) for x in g["bins_per_chr"]]
