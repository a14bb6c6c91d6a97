"""Host logic of the predict flow without a GPU (rows a15, f2): the flat result assembly against the
pinned NumPy restatement of the reference (oracle/np_oracle.py, itself checked against the reference's `tool_test`
golden in test_oracle_golden.py), `flatten`, and the batch plumbing of predict_control.predict_batch around a stand-in
for the C-ABI library (tests/fake_cabi.py: meaningless numbers of the right shape) -- every sample of a batch must get
exactly what it gets alone.  The arithmetic of the device calls is the business of the -m gpu tests."""
import os
import sys
import types

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import fake_cabi  # noqa: E402
from oracle import np_oracle as O  # noqa: E402
from wisecondorx_b200 import _lib, cbs, predict_control, predict_tools, synth  # noqa: E402


def _eq(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def test_flatten():
    f = predict_tools.flatten
    m = np.arange(24, dtype=np.float64).reshape(2, 12)
    parts = [m[0, 0:5], m[0, 5:5], m[0, 5:12], m[1, 0:4], m[1, 4:12]]
    v = f(parts)
    assert _eq(v, np.arange(24)) and np.shares_memory(v, m)
    parts[2][1] = -1.0  # writes through the per-chromosome views show in the flat vector
    assert v[6] == -1.0
    gap = [m[0, 0:5], m[0, 6:12]]  # not back to back: a copy
    v = f(gap)
    assert _eq(v, np.concatenate(gap)) and not np.shares_memory(v, m)
    assert not np.shares_memory(f([m[0, 0:5], m[0, 0:5]]), m)
    other = np.ones(3)
    assert _eq(f([m[0, 0:5], other[0:3]]), np.concatenate([m[0, 0:5], other]))
    assert _eq(f([np.arange(3), np.arange(2.0)]), [0, 1, 2, 0, 1])  # other dtypes are converted
    assert f([]).shape == (0,) and f([m[0, 2:2], m[0, 2:2]]).shape == (0,)
    one = np.arange(4.0)
    assert f([one]) is one
    i32 = np.arange(10, dtype=np.int32).reshape(2, 5)
    v = f([i32[0], i32[1]], np.int32)
    assert v.dtype == np.int32 and _eq(v, np.arange(10)) and np.shares_memory(v, i32)
    assert not np.shares_memory(f([m[0, 0:6:2], m[0, 6:12:2]]), m)  # strided views: a copy


def _fake_normalize(rng, n, nasty=True):
    r = 1.0 + 0.05 * rng.standard_normal(n)
    z = rng.standard_normal(n)
    w = rng.uniform(0.5, 2.0, n)
    nref = np.full(n, 300.0)
    if nasty:
        r[rng.random(n) < 0.05] = 0.0
        r[rng.random(n) < 0.01] = np.nan
        r[rng.random(n) < 0.01] = -1.0
        r[rng.random(n) < 0.01] = np.inf
        r[rng.random(n) < 0.02] = 1.0
        z[rng.random(n) < 0.01] = np.nan
        nref[rng.random(n) < 0.05] = 3.0
    return r, z, w, nref, float(rng.normal(0, 0.01)), float(rng.normal(0, 0.1))


@pytest.mark.parametrize("g,surplus,bad_weights", [("F", 0, False), ("M", 7, False), ("F", 4, True)])
def test_assemble_equals_pinned_restatement(tmp_path, g, surplus, bad_weights):
    """predict_control.assemble / assemble_batch (one native pass per sample over the bin axis, wcx_predict_assemble:
    host code, runs here) == np_oracle.assemble_results + log_trans + apply_blacklist (per
    key and per chromosome like the reference, pinned by test_oracle_golden.py) bit for bit: zero / negative / NaN /
    infinite ratios, bins with too few reference bins, more results than kept bins, non-numeric
    weights, a blacklist; one sample and rows of a batch."""
    rng = np.random.default_rng(11)
    ref, n_aut = fake_cabi.make_ref_file(binsize=500000, k=4, m=6)
    sfx = "." + g
    cnt = int(ref["mask" + sfx].sum())
    ct = int(ref["masked_bins_per_chr_cum" + sfx][21])
    n_gon = cnt - ct
    n_a = cnt - n_gon + surplus  # the autosomal pass saw `surplus` bins the gonosomal mask lost later
    aut = _fake_normalize(rng, n_a)
    gon = _fake_normalize(rng, n_gon)
    if bad_weights:
        gon = (gon[0], gon[1], np.full(n_gon, np.nan), gon[3], gon[4], gon[5])
    nr = np.zeros((n_a + n_gon, 6))
    bl = tmp_path / "bl.bed"
    bl.write_text("chr1\t1000000\t4200000\nchrX\t0\t900000\nY\t500000\t2500000\n7\t158000000\t999000000\n")
    args = types.SimpleNamespace(minrefbins=10, blacklist=str(bl))
    want, sizes = O.assemble_results(aut, gon, None, None, 10, ref["mask" + sfx], ref["bins_per_chr" + sfx])
    O.log_trans(want, aut[4])
    O.apply_blacklist(want, bl.read_text(), 500000)
    pos = np.arange(cnt, dtype=np.int32)
    pos[sizes[:cnt] < 10] = -1
    want_inflate = np.full(len(ref["mask" + sfx]), -1, dtype=np.int32)
    want_inflate[ref["mask" + sfx]] = pos

    r5 = lambda x, rows=(1, 3, 4): np.stack([x if i in rows else np.full_like(x, 7.0) for i in range(5)])  # noqa: E731
    gon3 = lambda x: np.stack([x, x, x])  # noqa: E731
    batch = predict_control.assemble_batch(args, (r5(aut[0]), r5(aut[1]), aut[2], r5(aut[3]), [9, aut[4], 9, aut[4], aut[4]],
                                                  [9, aut[5], 9, aut[5], aut[5]]), [3, 1, 4],
                                           (gon3(gon[0]), gon3(gon[1]), gon[2], gon3(gon[3])), nr, ref, g, [g] * 3, [123] * 3)
    for rem, got in [predict_control.assemble(args, aut, gon, nr, ref, g, g, 123)] + batch:
        assert rem["ref_gender"] == g and rem["n_reads"] == 123 and rem["binsize"] == 500000
        for key in ("results_r", "results_z", "results_w"):
            assert len(got[key]) == len(want[key]) == (23 if g == "F" else 24)
            for c in range(len(want[key])):
                assert _eq(got[key][c], want[key][c]), (key, c)
        assert got["results_nr"]["dense"] is nr and _eq(got["results_nr"]["inflate"], want_inflate)
        if bad_weights:
            kept = np.concatenate(got["results_w"]) != 0
            assert kept.any() and np.all(np.concatenate(got["results_w"])[kept] == 1.0)
    with pytest.raises(IndexError):  # fewer results than kept bins: the reference's inflate loop runs off its list
        predict_control.assemble(args, tuple(a[:-20] if isinstance(a, np.ndarray) else a for a in aut), gon, nr, ref, g, g, 1)


@pytest.fixture
def fake_library(monkeypatch):
    lib = fake_cabi.FakeLib(nasty=True)
    monkeypatch.setattr(_lib, "_lib", lib)
    monkeypatch.setattr(_lib, "_default_ctx", {})
    monkeypatch.setattr(predict_tools, "_engines", {})
    return lib


def test_predict_batch_equals_single_samples(fake_library, tmp_path):
    """predict_control.predict_batch: one normalize call per reference set, ONE CBS call for every chromosome of every
    sample, one z-score call per reference gender -- each sample's results (per-bin vectors, null-ratio map, segments,
    z-scores) are exactly those of a batch of one, for mixed genders, a blacklist and series with long runs without data."""
    binsize = 200000
    ref, _ = fake_cabi.make_ref_file(binsize=binsize, k=4, m=6)
    samples, genders = synth.make_samples(7, binsize, seed=2, depth=2e5)
    bl = tmp_path / "bl.bed"
    bl.write_text("chr3\t1000000\t9200000\nchrX\t0\t1900000\n")
    args = types.SimpleNamespace(maskrepeats=5, minrefbins=150, alpha=1e-4, seed=1, gender=None, blacklist=str(bl), zscore=5, beta=None)
    eng = predict_tools.PredictEngine(0)
    batch = predict_control.predict_batch(args, samples, [binsize] * len(samples), ref, eng)
    assert {rem["ref_gender"] for rem, _ in batch} == {"F", "M"}
    n_split = 0
    for i, s in enumerate(samples):
        rem1, res1 = predict_control.predict_batch(args, [s], [binsize], ref, eng)[0]
        rem, res = batch[i]
        assert (rem["gender"], rem["ref_gender"], rem["n_reads"]) == (rem1["gender"], rem1["ref_gender"], rem1["n_reads"])
        for key in ("results_r", "results_z", "results_w"):
            assert len(res[key]) == len(res1[key])
            assert all(_eq(a, b) for a, b in zip(res[key], res1[key])), key
        assert _eq(res["results_nr"]["inflate"], res1["results_nr"]["inflate"])
        assert len(res["results_c"]) == len(res1["results_c"]) > 0
        for a, b in zip(res["results_c"], res1["results_c"]):
            assert a[:3] == b[:3] and _eq(a[3], b[3]) and _eq(a[4], b[4]), (i, a, b)
        per_chr = np.bincount([sg[0] for sg in res["results_c"]], minlength=24)
        n_split += int((per_chr > 1).sum())
        # the per-chromosome vectors of a sample are views of one row (no copies on the way to the device calls)
        flat = predict_tools.flatten(res["results_r"])
        assert np.shares_memory(flat, res["results_r"][0]) and len(flat) == len(rem["mask"])
    assert n_split > 5


def test_cbs_batch_host_side_equals_oracle_postprocessing(fake_library):
    """cbs.cbs_segments_batch around the stand-in segmenter == the oracle's CBS.R restatement around the SAME
    segmenter, sample by sample (series with long NA runs, dropped chromosomes, weight 0)."""
    import ctypes
    from oracle import cbs_oracle as C
    rng = np.random.default_rng(3)
    batch = []
    for i in range(5):
        per = [int(x) for x in rng.integers(40, 900, 24)]
        rr = [rng.normal(0, 0.1, n) for n in per]
        ww = [rng.uniform(0.5, 2, n) for n in per]
        for r in rr:
            r[rng.random(len(r)) < 0.1] = 0
            a = int(rng.integers(0, len(r) - 1))
            r[a:a + int(rng.integers(1, 60))] = 0
        rr[i][:] = 0
        ww[2][::3] = 0
        batch.append((rr, ww, "M" if i % 2 else "F"))
    got = cbs.cbs_segments_batch(batch, 1e-4, 100000.0, seed=1, nperm=100)

    def segmenter(yy, wv, c):  # the stand-in's rule through its C-ABI signature
        n = len(yy)
        y = np.ascontiguousarray(yy, dtype=np.float64)
        off = np.array([0, n], dtype=np.int64)
        ids = np.array([c], dtype=np.int32)
        ends, nseg = np.zeros(n, dtype=np.int32), np.zeros(1, dtype=np.int32)
        p = lambda a: ctypes.c_void_p(a.ctypes.data)
        fake_library.wcx_cbs_segment(None, p(y), p(y), p(off), 1, p(ids), 1e-4, 100, 1, p(ends), p(nseg))
        return [int(x) for x in ends[:nseg[0]]]

    for (rr, ww, g), segs in zip(batch, got):
        want = [[d["chr"] - 1, d["s"], d["e"], d["r"]] for d in C.cbs_r(rr, ww, g, 1e-4, 100000.0, segmenter=segmenter)]
        assert len(segs) == len(want) > 20
        for a, b in zip(segs, want):
            assert a[:3] == b[:3] and _eq(a[3], b[3]), (a, b)


def test_segment_batch_series(fake_library):
    """predict_control.segment_batch hands the device call the same (sample, chromosome) series as the plain
    per-series construction: bins without a finite non-zero log ratio or with too few reference bins dropped, an empty
    chromosome, a chromosome without data."""
    rng = np.random.default_rng(0)
    b, n = 5, 4000
    offs = np.array([0, 700, 700, 1500, 2600, 4000])
    r = 1 + 0.05 * rng.standard_normal((b, n))
    for p, v in ((0.05, 0.0), (0.01, np.nan), (0.01, -1.0), (0.02, 1.0)):
        r[rng.random((b, n)) < p] = v
    r[2, 700:1500] = 0
    nref = np.full((b, n), 300.0)
    nref[rng.random((b, n)) < 0.05] = 10
    w = rng.uniform(0.5, 2, n)
    m_lr = rng.normal(0, 0.01, b)
    series = []
    for i in range(b):
        with np.errstate(all="ignore"):
            lr = np.log2(r[i]) - m_lr[i]
        ok = np.isfinite(lr) & (nref[i] >= 150) & (lr != 0)
        for c in range(len(offs) - 1):
            m = ok[offs[c]:offs[c + 1]]
            series.append((lr[offs[c]:offs[c + 1]][m], w[offs[c]:offs[c + 1]][m]))
    want = cbs.segment_series(series, [i % 5 for i in range(len(series))], alpha=1e-4, nperm=100, seed=3)
    got = predict_control.segment_batch(r, w, nref, m_lr, offs, nperm=100, seed=3)
    assert len(got) == len(want) == 25 and sum(len(x) for x in want) > 30
    assert all(np.array_equal(a, c) for a, c in zip(got, want))


def test_predict_command_line_host_flow(fake_library, tmp_path):
    """`WisecondorX predict --bed` end to end around the stand-in library: reference and sample files read, genders,
    batch-of-one flow, tables written -- the plumbing between the command line and the device calls (the numbers in the
    tables are the stand-in's)."""
    from wisecondorx_b200 import main as wmain
    binsize = 500000
    ref, _ = fake_cabi.make_ref_file(binsize=binsize, k=4, m=6)
    np.savez(tmp_path / "ref.npz", **ref)
    samples, _ = synth.make_samples(2, binsize, seed=9, depth=3e5)
    np.savez_compressed(tmp_path / "s.npz", binsize=binsize, sample=samples[1], quality={})
    parser = wmain.build_parser()
    a = parser.parse_args(["predict", str(tmp_path / "s.npz"), str(tmp_path / "ref.npz"), str(tmp_path / "out"), "--bed", "--seed", "3",
                           "--minrefbins", "150"])
    res = a.func(a)
    assert len(res["results_c"]) > 0 and set(res["timings"]) >= {"load_reference", "normalize_and_assemble", "cbs_and_segment_z", "write_tables"}
    bins = (tmp_path / "out_bins.bed").read_text().splitlines()
    assert bins[0].split("\t")[:4] == ["chr", "start", "end", "id"] and len(bins) == 1 + int(np.sum(ref["bins_per_chr.M"]))
    segs = (tmp_path / "out_segments.bed").read_text().splitlines()
    assert len(segs) == 1 + len(res["results_c"])
    for name in ("out_aberrations.bed", "out_statistics.txt"):
        assert (tmp_path / name).stat().st_size > 0


def test_newref_command_line_host_flow(fake_library, tmp_path):
    """`WisecondorX newref` end to end around the stand-in library: sample files read, one stacked count matrix for the
    Y fractions / masks / passes, the A, F and M passes with the PCA-distance filter and its redo, results copied out and
    deflated in the background, merge, QC -- then `predict` against the file it wrote.  The plumbing, not the numbers."""
    from wisecondorx_b200 import main as wmain, npz_io
    binsize = 1_000_000
    samples, genders = synth.make_samples(15, binsize, seed=6, depth=1e6)
    paths = []
    for i, s in enumerate(samples[:14]):
        paths.append(str(tmp_path / ("s%02d.npz" % i)))
        np.savez_compressed(paths[-1], binsize=binsize, sample=s, quality={})
    parser = wmain.build_parser()
    ref_path = str(tmp_path / "ref.npz")
    a = parser.parse_args(["newref"] + paths + [ref_path, "--binsize", str(binsize), "--yfrac", "0.006", "--refsize", "10", "--cpus", "2"])
    timings = a.func(a)
    assert {"prep.A", "prep.F", "prep.M", "get_reference.M", "write_reference", "qc_reference"} <= set(timings)
    ref = npz_io.load_npz(ref_path)
    plain = np.load(ref_path, allow_pickle=True)
    assert set(plain.files) == set(ref) and bool(ref["has_female"]) and bool(ref["has_male"]) and not bool(ref["is_nipt"])
    for sfx, nchr, ns in (("", 22, 14), (".F", 23, 7), (".M", 24, 7)):
        n = int(np.sum(ref["mask" + sfx]))
        assert len(ref["bins_per_chr" + sfx]) == nchr and int(ref["masked_bins_per_chr_cum" + sfx][-1]) == n
        assert ref["indexes" + sfx].shape == (n, 10) and ref["indexes" + sfx].dtype == np.int32
        assert ref["distances" + sfx].shape == (n, 10) and ref["null_ratios" + sfx].shape == (n, ns)
        assert ref["pca_components" + sfx].shape == (5, n) and ref["pca_mean" + sfx].shape == (n,)
        assert np.array_equal(plain["null_ratios" + sfx], ref["null_ratios" + sfx])
    # the in-place mask edit leaks from pass to pass: bins the filter removed stay removed
    la, lf = len(ref["mask"]), len(ref["mask.F"])
    assert not np.any(ref["mask.F"][:la] & ~ref["mask"]) and not np.any(ref["mask.M"][:lf] & ~ref["mask.F"])
    assert int(np.sum(ref["mask.M"][:la])) < int(np.sum(ref["mask"]))  # the redo ran in the later passes too
    # `newref --gpus 3`: every part cut into one row range per device, one context + host thread per device
    import random
    ref3_path = str(tmp_path / "ref3.npz")
    for pth, extra in ((ref_path, []), (ref3_path, ["--gpus", "3"])):
        random.seed(11)
        a = parser.parse_args(["newref"] + paths + [pth, "--binsize", str(binsize), "--yfrac", "0.006", "--refsize", "10", "--cpus", "2"] + extra)
        a.func(a)
    one, three = npz_io.load_npz(ref_path), npz_io.load_npz(ref3_path)
    assert set(one) == set(three)
    for key in one:
        assert np.array_equal(np.asarray(one[key]), np.asarray(three[key]), equal_nan=True), key
    np.savez_compressed(tmp_path / "t.npz", binsize=binsize, sample=samples[14], quality={})
    b = parser.parse_args(["predict", str(tmp_path / "t.npz"), ref_path, str(tmp_path / "out"), "--bed", "--minrefbins", "5"])
    res = b.func(b)
    assert len(res["results_c"]) > 0 and (tmp_path / "out_bins.bed").stat().st_size > 0
