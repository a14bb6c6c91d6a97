"""f2 / f3: the vectorised host functions of the drop-in package against the LIVE reference functions they mirror
(scale_sample, gender model, get_mask, get_post_processed_result / inflate_results, log_trans, apply_blacklist, the
bins / segments / aberrations writers, segment statistics, ref QC).  Pure host code: runs where /root/reference exists
(the build container); on the GPU box the committed goldens (tests/test_assembly_gpu.py) cover the same functions."""
import copy
import logging
import os
import types

import numpy as np
import pytest

from oracle import ref_loader
from wisecondorx_b200 import main as wcx_main, overall_tools, predict_control, predict_output, predict_tools, ref_qc, synth

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="live reference not present (GPU box)")


@pytest.fixture(scope="module")
def R():
    return ref_loader.load()


@pytest.mark.parametrize("frm,to", [(5000, 15000), (5000, 100000), (100000, 100000), (50000, 1000000)])
def test_scale_sample(R, frm, to):
    rng = np.random.default_rng(frm + to)
    sample = {str(c): rng.poisson(30, int(rng.integers(1, 700))).astype(np.int32) for c in range(1, 25)}
    sample["7"] = np.zeros(0, dtype=np.int32) if frm != to else sample["7"]
    want = R.overall_tools.scale_sample(copy.deepcopy(sample), frm, to)
    got = overall_tools.scale_sample(copy.deepcopy(sample), frm, to)
    assert sorted(got) == sorted(want)
    for k in want:
        assert np.array_equal(got[k], want[k]), k
        assert np.asarray(got[k]).dtype == np.asarray(want[k]).dtype, k


def test_scale_sample_rejects_like_reference(R):
    sample = {"1": np.arange(10, dtype=np.int32)}
    for fn in (R.overall_tools.scale_sample, overall_tools.scale_sample):
        with pytest.raises(SystemExit):
            fn(sample, 5000, 7000)


def test_get_mask_and_gender_model(R):
    samples, genders = synth.make_samples(16, 2_000_000, seed=12, depth=5e6)
    arr = np.array(samples)
    want_mask, want_bpc = R.newref_tools.get_mask(arr)
    got_mask, got_bpc = wcx_main.get_mask(arr)
    assert np.array_equal(got_mask, want_mask) and list(got_bpc) == list(want_bpc)
    # masks of sample subsets out of ONE stacked count matrix (tool_newref: all samples, then the females, then the males)
    from wisecondorx_b200 import newref_tools
    counts_all = newref_tools.stack_counts(list(arr), range(1, 25))
    g = np.array(genders)
    for sub in (np.ones(len(arr), bool), g == "F", g == "M", np.arange(len(arr)) % 3 == 0):
        want_sub, want_sub_bpc = R.newref_tools.get_mask(arr[sub])
        got_sub, got_sub_bpc = wcx_main.get_mask(arr[sub], counts_all, np.flatnonzero(sub))
        assert np.array_equal(wcx_main.get_mask(arr[sub], counts_all, np.flatnonzero(sub), newref_tools.column_totals(counts_all))[0], want_sub)
        assert np.array_equal(got_sub, want_sub) and list(got_sub_bpc) == list(want_sub_bpc)
        # the count matrix of a gonosomal pass: row prefix + column subset of the stacked matrix (threaded copy)
        rows = counts_all.shape[0] - 7
        taken = newref_tools.take_columns(counts_all, rows, np.flatnonzero(sub))
        assert taken.flags.c_contiguous and taken.dtype == counts_all.dtype
        assert np.array_equal(taken, counts_all[:rows][:, sub])
        assert np.array_equal(taken, newref_tools.stack_counts(list(arr[sub]), range(1, 25))[:rows])
    short = [dict(smp) for smp in arr[:5]]  # a subset with fewer bins than the stacked matrix: falls back to its own stack
    for smp in short:
        smp["7"] = smp["7"][:-3]
    want_sub, _ = R.newref_tools.get_mask(np.array(short))
    got_sub, _ = wcx_main.get_mask(np.array(short), counts_all, np.arange(5))
    assert np.array_equal(got_sub, want_sub)
    args = types.SimpleNamespace(yfrac=0.006, plotyfrac=None)
    want_g, want_cut = R.newref_tools.train_gender_model(args, arr)
    got_g, got_cut = wcx_main.train_gender_model(args, arr)
    assert got_g == want_g and got_cut == want_cut
    for s in samples[:4]:
        assert wcx_main.predict_gender(s, 0.006) == R.predict_tools.predict_gender(s, 0.006)


def test_gender_model_gmm(R):
    """Without --yfrac: GaussianMixture + first local minimum (newref_tools.py:21-68).  Needs a bimodal Y fraction."""
    samples, genders = synth.make_samples(40, 5_000_000, seed=8, depth=5e6)
    arr = np.array(samples)
    args = types.SimpleNamespace(yfrac=None, plotyfrac=None)
    np.random.seed(1)
    try:
        want_g, want_cut = R.newref_tools.train_gender_model(args, arr)
    except IndexError:
        pytest.skip("the reference finds no local minimum on this synthetic Y-fraction distribution (SURVEY.md 8c)")
    np.random.seed(1)
    got_g, got_cut = wcx_main.train_gender_model(args, arr)
    assert got_g == want_g and np.isclose(got_cut, want_cut, rtol=1e-12)


def _fake_results(rng, bpc, zero_frac=0.1):
    res = {"results_r": [], "results_z": [], "results_w": []}
    for nb in bpc:
        r = np.exp2(rng.normal(0, 0.1, nb))
        r[rng.random(nb) < zero_frac] = 0
        r[rng.random(nb) < 0.02] = -1.0  # log2 of a negative ratio -> nan -> blanked
        res["results_r"].append(r)
        res["results_z"].append(rng.normal(0, 1, nb))
        res["results_w"].append(rng.uniform(0.5, 2, nb))
    return res


def test_post_processing_chain(R, tmp_path):
    rng = np.random.default_rng(2)
    bpc = [int(x) for x in rng.integers(5, 60, 24)]
    total = sum(bpc)
    mask = rng.random(total) > 0.15
    nm = int(mask.sum())
    ref_sizes = rng.integers(0, 300, nm)
    vals = rng.normal(1, 0.1, nm)
    rem = {"mask": mask, "bins_per_chr": bpc, "binsize": 100000}
    args = types.SimpleNamespace(minrefbins=150)
    want = R.predict_control.get_post_processed_result(args, vals.copy(), ref_sizes, rem)
    got = wcx_main.get_post_processed_result(150, vals.copy(), ref_sizes, mask, bpc)
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert np.array_equal(np.asarray(a, dtype=float), np.asarray(b, dtype=float))
    assert np.array_equal(predict_control.inflate_results(vals, mask), np.array(R.predict_tools.inflate_results(vals, rem), dtype=float))
    # log_trans
    res_w = _fake_results(rng, bpc)
    res_g = copy.deepcopy(res_w)
    want_res = {k: [x.copy() for x in v] for k, v in res_w.items()}
    with np.errstate(all="ignore"):
        R.predict_tools.log_trans(want_res, 0.0123)
    wcx_main.log_trans(res_g, 0.0123)
    for key in ("results_r", "results_z", "results_w"):
        for c in range(24):
            np.testing.assert_array_equal(np.asarray(res_g[key][c], dtype=float), np.asarray(want_res[key][c], dtype=float), err_msg=key)
    # apply_blacklist (chr prefix, X / Y names, out-of-range and clipped intervals, chrY skipped with 23 chromosomes)
    bl = tmp_path / "bl.bed"
    bl.write_text("chr1\t150000\t420000\n2\t0\t99999\nX\t200000\t100000000\nY\t0\t300000\nchr5\t-5\t10\n")
    for nchr in (24, 23):
        a = {k: [np.array(x, dtype=float) for x in v[:nchr]] for k, v in res_g.items()}
        b = {k: [np.array(x, dtype=float).tolist() for x in v[:nchr]] for k, v in res_g.items()}
        rem2 = {"args": types.SimpleNamespace(blacklist=str(bl)), "binsize": 100000}
        R.predict_tools.apply_blacklist(rem2, b)
        wcx_main.apply_blacklist(str(bl), 100000, a)
        for key in a:
            for c in range(nchr):
                np.testing.assert_array_equal(a[key][c], np.asarray(b[key][c], dtype=float))


def test_raw_vector_equals_reference_coverage_normalisation(R):
    """host half of coverage_normalize_and_mask (predict_tools.py:35-44): pad / truncate per chromosome."""
    rng = np.random.default_rng(4)
    bpc = [int(x) for x in rng.integers(5, 40, 24)]
    sample = {str(c + 1): rng.poisson(20, bpc[c] + int(rng.integers(-3, 4))).astype(np.int32) for c in range(24)}
    mask = rng.random(sum(bpc)) > 0.2
    rem = {"bins_per_chr": bpc, "mask": mask}
    ref_file = {"bins_per_chr": bpc, "mask": mask}
    want = R.predict_tools.coverage_normalize_and_mask(sample, ref_file, "")
    raw = predict_tools.raw_vector(sample, bpc)
    np.testing.assert_allclose((raw / raw.sum())[mask], want, rtol=1e-15)


def test_table_writers(R, tmp_path):
    rng = np.random.default_rng(6)
    bpc = [int(x) for x in rng.integers(4, 30, 24)]
    res = _fake_results(rng, bpc, 0.2)
    for c in range(24):
        res["results_r"][c] = np.where(res["results_r"][c] > 0, np.log2(np.abs(res["results_r"][c]) + 1e-9), 0.0)
    segs = []
    for c in range(24):
        h = bpc[c] // 2
        segs.append([c, 0, h, float(rng.normal(0, 4)), float(rng.normal(0, 0.2))])
        segs.append([c, h, bpc[c], "nan" if c % 5 == 0 else float(rng.normal(0, 8)), float(rng.normal(0, 0.2))])
    res["results_c"] = segs
    for beta, gender in ((None, "F"), (0.4, "M")):
        outs = []
        for tag, mod in (("ref", R.predict_output), ("got", predict_output)):
            outid = str(tmp_path / f"{tag}{gender}")
            rem = {"args": types.SimpleNamespace(outid=outid, beta=beta, zscore=5, regions=None), "binsize": 100000,
                   "ref_gender": gender, "gender": gender, "n_reads": 123, "bins_per_chr": bpc}
            r2 = {k: ([np.asarray(x).tolist() for x in v] if tag == "ref" and k != "results_c" else copy.deepcopy(v)) for k, v in res.items()}
            mod._generate_bins_bed(rem, r2)
            mod._generate_segments_and_aberrations_bed(rem, r2)
            outs.append(outid)
        for sfx in ("_bins.bed", "_segments.bed", "_aberrations.bed"):
            assert open(outs[0] + sfx).read() == open(outs[1] + sfx).read(), sfx
    # segment statistics helpers
    lists = [np.asarray(x).tolist() for x in res["results_r"]]
    assert overall_tools.get_median_segment_variance(segs, res["results_r"]) == R.overall_tools.get_median_segment_variance(segs, lists)
    num = [sg for sg in segs if not isinstance(sg[3], str)]  # both raise TypeError on a "nan" z-score
    assert overall_tools.get_cpa(num, 100000) == R.overall_tools.get_cpa(num, 100000)


def test_regions_bed_autosomes(R, tmp_path):
    """_generate_regions_bed (predict_output.py:86-137) on autosomal regions (the reference raises on X / Y, A.8)."""
    rng = np.random.default_rng(9)
    bpc = [int(x) for x in rng.integers(10, 30, 24)]
    res = _fake_results(rng, bpc, 0.1)
    reg = tmp_path / "regions.bed"
    reg.write_text("chr1\t0\t450000\tA\n2\t300000\t99999999\tB\n\nchr3\t500000\t100000\tbad\n")
    outs = []
    for tag, mod in (("ref", R.predict_output), ("got", predict_output)):
        outid = str(tmp_path / tag)
        rem = {"args": types.SimpleNamespace(outid=outid, regions=str(reg)), "binsize": 100000, "bins_per_chr": bpc}
        mod._generate_regions_bed(rem, {k: [np.asarray(x) for x in v] for k, v in res.items()})
        outs.append(outid + "_regions.bed")
    assert open(outs[0]).read() == open(outs[1]).read()


def test_ref_qc_matches_reference(R, tmp_path, caplog):
    rng = np.random.default_rng(11)
    ref = {"binsize": 100000, "is_nipt": False, "trained_cutoff": 0.005, "has_female": True, "has_male": True}
    for sfx, nchr in ((".F", 23), (".M", 24), ("", 22)):
        per = rng.integers(20, 60, nchr)
        n = int(per.sum())
        ref["bins_per_chr" + sfx] = per
        ref["masked_bins_per_chr" + sfx] = per
        ref["masked_bins_per_chr_cum" + sfx] = np.cumsum(per)
        ref["indexes" + sfx] = rng.integers(0, n, (n, 160 if sfx != ".F" else 100)).astype(np.int32)
        d = np.sort(rng.random((n, ref["indexes" + sfx].shape[1])) * (3 if sfx == ".M" else 1), axis=1)
        if sfx == ".M":
            d[int(np.cumsum(per)[22]):] *= 9  # poor chrY
        ref["distances" + sfx] = d
    path = str(tmp_path / "ref.npz")
    np.savez_compressed(path, **ref)
    with caplog.at_level(logging.INFO):
        caplog.clear()
        want = R.ref_qc.qc_reference(path)
        want_msgs = [r.getMessage() for r in caplog.records]
        caplog.clear()
        got = ref_qc.qc_reference(path)
        got_msgs = [r.getMessage() for r in caplog.records]
        caplog.clear()
        got2 = ref_qc.qc_reference(path, ref)
    assert got == want == got2
    assert got_msgs == want_msgs
    for sfx in (".F", ".M"):
        a = ref_qc.compute_metrics(ref, sfx)
        b = R.ref_qc._compute_metrics(ref, sfx)
        assert a == b


def test_stack_counts_native_equals_column_fill():
    """newref_tools.stack_counts (wcx_host_stack_counts: blocked transposition on host threads) == the reference's
    column-by-column fill (newref_tools.py:81-92): ragged samples, empty chromosomes, other integer / float dtypes,
    chromosome subsets; column_totals; take_columns."""
    from wisecondorx_b200 import newref_tools
    rng = np.random.default_rng(0)

    def fill(samples, chrs):
        chrs = list(chrs)
        lens = [max(len(s[str(c)]) for s in samples) for c in chrs]
        offs = np.concatenate([[0], np.cumsum(lens)]).astype(int)
        out = np.zeros((offs[-1], len(samples)), dtype=np.int32)
        for i, s in enumerate(samples):
            for c, o in zip(chrs, offs[:-1]):
                out[o:o + len(s[str(c)]), i] = s[str(c)]
        return out

    for trial in range(12):
        samples = []
        for i in range(int(rng.integers(1, 40))):
            s = {}
            for c in range(1, 25):
                n = int(rng.integers(0, 60)) if trial % 3 else 37
                n = 0 if (trial % 4 == 0 and c == 5) else n
                dt = [np.int32, np.int64, np.float64, np.uint16][int(rng.integers(0, 4))] if trial % 2 else np.int32
                s[str(c)] = rng.integers(0, 1000, n).astype(dt)
            samples.append(s)
        for chrs in (range(1, 25), range(1, 23), [3], [5, 6]):
            got = newref_tools.stack_counts(samples, chrs)
            assert got.dtype == np.int32 and got.flags.c_contiguous and np.array_equal(got, fill(samples, chrs))
        assert np.array_equal(newref_tools.column_totals(got), got.sum(0, dtype=np.int64))
    big, _ = synth.make_samples(21, 200000, seed=4)
    got = newref_tools.stack_counts(big, range(1, 25))
    assert np.array_equal(got, fill(big, range(1, 25)))
    assert np.array_equal(newref_tools.column_totals(got), got.sum(0, dtype=np.int64))


@pytest.mark.parametrize("s", [1, 5, 8, 9, 64, 127, 128, 129, 255, 256, 257, 500, 777, 1100, 2049])
def test_bin_sums_native_equals_numpy_row_sums(s):
    """newref_tools.bin_sums (wcx_host_bin_sums) returns the float64 VALUE of np.sum(counts / col_sum, 1) -- NumPy's
    pairwise order replayed for every row length class -- for all columns and for a column subset."""
    from wisecondorx_b200 import newref_tools
    rng = np.random.default_rng(s)
    counts = rng.poisson(30.0, (301, s)).astype(np.int32)
    counts[rng.random(counts.shape) < 0.1] = 0
    counts[7] = 0
    col = counts.sum(0, dtype=np.int64).astype(float)
    want = np.sum(counts.astype(float) / col, 1)
    assert np.array_equal(newref_tools.bin_sums(counts, col), want, equal_nan=True)
    cols = np.flatnonzero(rng.random(s) < 0.5)
    if len(cols):
        # (the subset's own C-ordered matrix, as the reference builds it: a fancy column index alone is F-ordered and
        # NumPy then adds column by column)
        want = np.sum(np.ascontiguousarray(counts[:, cols]).astype(float) / col[cols], 1)
        assert np.array_equal(newref_tools.bin_sums(counts, col[cols], cols), want, equal_nan=True)


def test_stacked_counts_gender_path(R):
    """main.stacked_counts: Y fractions, genders and the gender-corrected count matrix from ONE stacked matrix == the
    reference's per-sample route (train_gender_model on the dicts, gender_correct, stacking the corrected samples)."""
    from wisecondorx_b200 import newref_tools
    samples, _ = synth.make_samples(18, 1_000_000, seed=21, depth=2e6)
    st = wcx_main.stacked_counts(np.array(samples))
    assert st.y_fractions is not None
    assert st.y_fractions.tolist() == [wcx_main._y_fraction(s) for s in samples]
    args = types.SimpleNamespace(yfrac=0.006, plotyfrac=None)
    want_g, want_cut = R.newref_tools.train_gender_model(args, np.array(samples))
    got_g, got_cut = wcx_main.train_gender_model(args, np.array(samples), st.y_fractions)
    assert got_g == want_g and got_cut == want_cut and {"F", "M"} <= set(got_g)
    corrected = [R.overall_tools.gender_correct(dict(s), g) for s, g in zip(samples, want_g)]
    st.gender_correct(got_g)
    want = newref_tools.stack_counts(corrected, range(1, 25))
    assert np.array_equal(st.counts, want)
    assert np.array_equal(st.totals, want.sum(0, dtype=np.int64))
    want_mask, _ = R.newref_tools.get_mask(np.array(corrected))
    assert np.array_equal(wcx_main.get_mask(np.array(corrected), st.counts, None, st.totals)[0], want_mask)
    # a sample with an extra key or without reads: the per-sample route decides
    odd = [dict(s) for s in samples]
    odd[0]["extra"] = np.ones(3, dtype=np.int32)
    assert wcx_main.stacked_counts(np.array(odd)).y_fractions is None
    empty = [dict(s) for s in samples]
    empty[1] = {k: np.zeros_like(v) for k, v in empty[1].items()}
    assert wcx_main.stacked_counts(np.array(empty)).y_fractions is None


def test_ref_qc_row_blocks_equal_whole_array():
    """ref_qc.compute_per_bin_stats on row blocks (thread pool, above 32 768 bins) == the whole-array reductions."""
    from wisecondorx_b200 import ref_qc
    rng = np.random.default_rng(2)
    d = rng.random((70001, 37)) * 3
    idx = np.zeros(d.shape, dtype=np.int32)
    mean_d, max_d, n_refs = ref_qc.compute_per_bin_stats(idx, d)
    assert np.array_equal(mean_d, np.mean(d, axis=1)) and np.array_equal(max_d, np.max(d, axis=1)) and np.all(n_refs == 37)
    mean_d, max_d, _ = ref_qc.compute_per_bin_stats(idx, d, need_max=False)
    assert max_d is None and np.array_equal(mean_d, np.mean(d, axis=1))


def test_native_float_text_equals_python_repr(tmp_path):
    """wcx_host_format_repr / wcx_host_format_bins (csrc/host_tables.cu): the text of a float64 is Python's repr -- what
    the reference's str(np.float64) prints -- for normal, tiny, huge, subnormal, integral and special values and for
    random bit patterns; the bins table equals the line-by-line Python formatting of the reference
    (predict_output.py:59-84)."""
    import ctypes

    from wisecondorx_b200 import _lib, predict_output
    L = _lib.load()
    rng = np.random.default_rng(0)
    sets = [rng.standard_normal(100000), rng.standard_normal(100000) * 0.05,
            2.0 ** rng.integers(-1074, 1024, 50000) * rng.random(50000),
            rng.integers(-10 ** 6, 10 ** 6, 20000).astype(float),
            10.0 ** rng.integers(-30, 30, 20000) * rng.integers(1, 1000, 20000),
            np.array([0.0, -0.0, 1.0, -1.0, 1e16, 9999999999999998.0, 1e-4, 9.999e-5, 1e-5, 1e22, 1e23, 5e-324,
                      1.7976931348623157e308, np.nan, np.inf, -np.inf, 0.1, 0.30000000000000004, 123456789012345678.0,
                      1e15, 1.5e16, 100.0, 1e100, 1e-100, 2.5e-5, 12345.678]),
            np.frombuffer(rng.bytes(8 * 100000), dtype=np.float64)]
    for x in sets:
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty(len(x) * 26, dtype=np.uint8)
        n = ctypes.c_int64()
        _lib.check(L.wcx_host_format_repr(ctypes.c_void_p(x.ctypes.data), len(x), ctypes.c_void_p(out.ctypes.data), out.nbytes, ctypes.byref(n)))
        got = out[:n.value].tobytes().decode("ascii").split("\n")[:-1]
        assert got == [repr(v) for v in x.tolist()]
    # the whole table: native fast path (float64 arrays) against the generic per-line path (lists)
    r = [rng.standard_normal(n) * 0.1 for n in (300, 1, 0, 57)]
    z = [rng.standard_normal(n) * 2 for n in (300, 1, 0, 57)]
    for a in r + z:
        a[rng.random(len(a)) < 0.2] = 0
    r[0][5], z[0][6] = np.nan, np.inf
    for binsize in (15000, np.int64(100000), 1):
        rem = {"binsize": binsize, "args": types.SimpleNamespace(outid=str(tmp_path / "a"))}
        predict_output._generate_bins_bed(rem, {"results_r": r, "results_z": z})
        rem["args"].outid = str(tmp_path / "b")
        predict_output._generate_bins_bed(rem, {"results_r": [list(a) for a in r], "results_z": [list(a) for a in z]})
        assert (tmp_path / "a_bins.bed").read_text() == (tmp_path / "b_bins.bed").read_text()
