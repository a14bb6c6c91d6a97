"""End-to-end GPU test of the drop-in CLI: `newref` on synthetic samples writes a reference .npz with
the reference's key layout / dtypes, `predict --bed` on it finds a planted gain and writes the
reference's four tables."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from wisecondorx_b200 import main as wcx_main, synth  # noqa: E402

BINSIZE = 1_000_000


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli")
    # 31 samples from one generative model (same bin profile / bias factors); the last one carries the CNV
    samples, genders = synth.make_samples(31, BINSIZE, seed=31, depth=6e6, cnv=[(30, 5, 40, 90, 1.5)])
    files = []
    for i, s in enumerate(samples[:30]):
        f = os.path.join(d, f"s{i}.npz")
        np.savez_compressed(f, binsize=BINSIZE, sample=s, quality={})
        files.append(f)
    tf = os.path.join(d, "test.npz")
    np.savez_compressed(tf, binsize=BINSIZE, sample=samples[30], quality={})
    return d, files, tf


def test_newref_then_predict(workdir):
    d, files, tf = workdir
    ref = os.path.join(d, "ref.npz")
    wcx_main.main(["newref"] + files + [ref, "--binsize", str(BINSIZE), "--refsize", "100", "--yfrac", "0.006", "--cpus", "3"])
    r = np.load(ref, allow_pickle=True)
    for sfx in ("", ".F", ".M"):
        for key in ("binsize", "mask", "bins_per_chr", "masked_bins_per_chr", "masked_bins_per_chr_cum", "pca_components",
                    "pca_mean", "indexes", "distances", "null_ratios"):
            assert key + sfx in r.files, key + sfx
    for key in ("has_female", "has_male", "is_nipt", "trained_cutoff"):
        assert key in r.files
    n = int(r["masked_bins_per_chr_cum"][-1])
    assert r["indexes"].dtype == np.int32 and r["indexes"].shape == (n, 100)
    assert r["distances"].dtype == np.float64 and r["null_ratios"].shape == (n, 30)
    assert r["pca_components"].shape == (5, n) and len(r["bins_per_chr.M"]) == 24
    assert (np.diff(r["distances"], axis=1) >= 0).all()
    # gonosomal references carry placeholder rows for the autosomes (newref_tools.py:186-191)
    nf = int(r["masked_bins_per_chr_cum.F"][21])
    assert (r["indexes.F"][:nf] == 0).all() and (r["distances.F"][:nf] == 1.0).all()

    outid = os.path.join(d, "out")
    res = wcx_main.main(["predict", tf, ref, outid, "--bed", "--minrefbins", "50", "--seed", "1"])
    for sfx in ("_bins.bed", "_segments.bed", "_aberrations.bed", "_statistics.txt"):
        assert os.path.exists(outid + sfx)
    ab = [l.split("\t") for l in open(outid + "_aberrations.bed").read().splitlines()[1:]]
    gains = [a for a in ab if a[0] == "5" and a[5] == "gain"]
    assert gains, ab
    start, end = int(gains[0][1]), int(gains[0][2])
    assert abs(start - (40 * BINSIZE + 1)) <= 3 * BINSIZE and abs(end - 90 * BINSIZE) <= 3 * BINSIZE
    assert 0.4 < float(gains[0][3]) < 0.7  # log2(1.5) = 0.585
    bins = open(outid + "_bins.bed").read().splitlines()
    assert bins[0] == "chr\tstart\tend\tid\tratio\tzscore" and len(bins) == 1 + int(np.sum(r["bins_per_chr.F"]))


def test_newref_multi_device_equals_single(workdir):
    """`newref --gpus N`: the target bins of every part are cut into one row range per device, every device context gets
    its own copy of the prepared matrix (wcx_newref_load x_on_device = 3) and writes its rows into the shared result
    arrays from its own host thread.  Three contexts on the one GPU of the test box exercise exactly that path; the result
    must equal the single-context run bit for bit (same null-sample draws: random is seeded)."""
    import random
    from wisecondorx_b200 import main as wmain, newref_control
    samples, genders = synth.make_samples(24, 2_000_000, seed=77, depth=6e6)
    arr = np.array(samples)
    mask, bpc = wmain.get_mask(arr)
    outs = []
    for devices, parts in (([0], 1), ([0, 0, 0], 1), ([0, 0], 3)):
        prep = newref_control.tool_newref_prep(list(arr), "A", mask.copy(), bpc, 0)
        random.seed(11)
        outs.append(newref_control.tool_newref_main(prep, 80, parts, 0, devices))
    a, b, c = outs
    assert np.array_equal(a["indexes"], b["indexes"]) and np.array_equal(a["distances"], b["distances"])
    assert np.array_equal(a["null_ratios"], b["null_ratios"], equal_nan=True)
    # three parts draw three sets of null samples: indexes / distances are unaffected
    assert np.array_equal(a["indexes"], c["indexes"]) and np.array_equal(a["distances"], c["distances"])
    assert c["null_ratios"].shape == a["null_ratios"].shape
