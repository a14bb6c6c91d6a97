"""a15 / f2: the drop-in `tool_test` (wisecondorx_b200/main.py) against the golden captured from the LIVE
reference's `tool_test` (main.py:145-300) run on the reference .npz the reference built itself, with the R bridge
stubbed on both sides by the same function (tests/golden/stub_cbs.py; R / DNAcopy are absent).
Covers the result assembly (main.py:242-271), get_post_processed_result, log_trans, apply_blacklist, get_z_score as
called from exec_cbs and from the statistics table, and the text of the four output tables."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from stub_cbs import stub_segments  # noqa: E402
from wisecondorx_b200 import cbs, main as wcx_main  # noqa: E402


@pytest.fixture(scope="module")
def gold(golden_dir):
    return (np.load(os.path.join(golden_dir, "newref_predict.npz"), allow_pickle=True),
            np.load(os.path.join(golden_dir, "tool_test.npz"), allow_pickle=True))


def _cmp_table(got_text, want_text, rtol, atol=1e-9):
    got, want = got_text.splitlines(), want_text.splitlines()
    assert len(got) == len(want)
    for gl, wl in zip(got, want):
        gf, wf = gl.split("\t"), wl.split("\t")
        assert len(gf) == len(wf), (gl, wl)
        for a, b in zip(gf, wf):
            if a == b:
                continue
            try:
                fa, fb = float(a), float(b)
            except ValueError:
                # trailing free-text lines of the statistics file: "label: number"
                la, _, na = a.rpartition(": ")
                lb, _, nb = b.rpartition(": ")
                assert la == lb, (gl, wl)
                fa, fb = float(na), float(nb)
            assert np.isclose(fa, fb, rtol=rtol, atol=atol, equal_nan=True), (gl, wl)


@pytest.mark.parametrize("si", [0, 1])
def test_tool_test_matches_reference(gold, tmp_path, monkeypatch, si):
    gpred, gtool = gold
    ref = {k[5:]: gpred[k] for k in gpred.files if k.startswith("ref__")}
    ref_path = str(tmp_path / "ref.npz")
    np.savez(ref_path, **ref)
    binsize = int(ref["binsize"])
    sample = {str(c): gpred[f"t{si}_sample_{c}"] for c in range(1, 25)}
    infile = str(tmp_path / "test.npz")
    np.savez_compressed(infile, binsize=binsize, sample=sample, quality={})
    bl = None
    if si == 0:
        bl = str(tmp_path / "blacklist.bed")
        open(bl, "w").write(str(gtool["blacklist_text"]))

    def stub(results_r, results_w, ref_gender, alpha, bs, seed=None, nperm=10000, ctx=None):
        return stub_segments(results_r, results_w, 24 if ref_gender == "M" else 23)

    monkeypatch.setattr(cbs, "cbs_segments", stub)
    outid = str(tmp_path / "out")
    argv = ["predict", infile, ref_path, outid, "--bed", "--minrefbins", "10", "--seed", "1"]
    if bl:
        argv += ["--blacklist", bl]
    parser = wcx_main.build_parser()
    args = parser.parse_args(argv)
    res = wcx_main.tool_test(args)
    meta = gtool[f"t{si}_meta"]
    for key, tol in (("results_r", 1e-9), ("results_z", 1e-6), ("results_w", 1e-12)):
        got = np.concatenate([np.asarray(x, dtype=float) for x in res[key]])
        want = gtool[f"t{si}_{key}"]
        assert np.array_equal(got == 0, want == 0), key
        np.testing.assert_allclose(got, want, rtol=tol, atol=1e-10, err_msg=key)
    want_c = gtool[f"t{si}_results_c"]
    got_c = np.array([[c[0], c[1], c[2], np.nan if isinstance(c[3], str) else c[3], c[4]] for c in res["results_c"]], dtype=float)
    assert got_c.shape == want_c.shape
    assert np.array_equal(got_c[:, :3], want_c[:, :3])
    np.testing.assert_allclose(got_c[:, 4], want_c[:, 4], rtol=1e-9)           # segment ratios
    np.testing.assert_allclose(got_c[:, 3], want_c[:, 3], rtol=1e-6, equal_nan=True)  # segment z-scores (get_z_score)
    assert os.path.exists(outid + "_bins.bed")
    _cmp_table(open(outid + "_bins.bed").read(), str(gtool[f"t{si}_bins.bed"]), 1e-6)
    _cmp_table(open(outid + "_segments.bed").read(), str(gtool[f"t{si}_segments.bed"]), 1e-6)
    _cmp_table(open(outid + "_aberrations.bed").read(), str(gtool[f"t{si}_aberrations.bed"]), 1e-6)
    _cmp_table(open(outid + "_statistics.txt").read(), str(gtool[f"t{si}_statistics.txt"]), 1e-5, atol=2e-5)
    assert str(meta[0]) in ("F", "M")


def test_predict_batch_equals_single_samples(gold):
    """predict_control.predict_batch (one normalize call per reference set, ONE CBS call and one z-score call per
    gender for the whole batch) returns for every sample exactly what it returns for that sample alone: segments,
    segment z-scores and the per-bin results."""
    import types
    from wisecondorx_b200 import predict_control, predict_tools
    gpred, _ = gold
    ref = {k[5:]: gpred[k] for k in gpred.files if k.startswith("ref__")}
    ref = {k: (v.item() if v.shape == () else v) for k, v in ref.items()}
    binsize = int(ref["binsize"])
    samples = [{str(c): gpred[f"t{si}_sample_{c}"] for c in range(1, 25)} for si in (0, 1, 0, 1, 1)]
    args = types.SimpleNamespace(maskrepeats=5, minrefbins=10, alpha=1e-4, seed=3, gender=None, blacklist=None, zscore=5, beta=None)
    eng = predict_tools.PredictEngine(0)
    batch = predict_control.predict_batch(args, samples, [binsize] * len(samples), ref, eng)
    assert len(batch) == len(samples)
    for s, (rem, res) in zip(samples, batch):
        rem1, res1 = predict_control.predict_batch(args, [s], [binsize], ref, eng)[0]
        assert rem1["ref_gender"] == rem["ref_gender"] and rem1["gender"] == rem["gender"]
        assert res1["results_c"] == res["results_c"]
        for key in ("results_r", "results_z", "results_w"):
            for a, b in zip(res1[key], res[key]):
                assert np.array_equal(a, b)
    assert {rem["ref_gender"] for rem, _ in batch} == {"F", "M"}
