"""Generates the golden vectors under tests/golden/ from the LIVE reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Every array is produced by calling the unmodified reference functions
(/root/reference/src/wisecondorx) through oracle/ref_loader.py with pinned seeds; the
fixtures travel to the GPU box, the reference does not.
"""
from __future__ import annotations

import argparse
import os
import random
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import ref_loader  # noqa: E402
from wisecondorx_b200 import synth  # noqa: E402


def golden_get_reference(R):
    """get_reference (newref_tools.py:155) on three cases: autosomal in parts, gonosomal,
    and a quantised/short case exercising ties, -1 / 1e10 fillers."""
    out = {}
    # case A: 22 chromosomes, parts 1/1 and 2/3
    per = [40, 35, 30, 28, 25, 22, 20, 18, 17, 16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 5]
    x, per, cum = synth.make_corrected_matrix(per, 12, seed=1)
    out["A_x"], out["A_per"], out["A_cum"] = x, per, cum
    for tag, (part, parts) in {"p11": (1, 1), "p23": (2, 3)}.items():
        random.seed(7)
        i, d, n = R.newref_tools.get_reference(x, per, cum, 30, part, parts)
        random.seed(7)
        out[f"A_{tag}_ids"] = np.array(random.sample(range(x.shape[1]), min(x.shape[1], 100)))
        out[f"A_{tag}_idx"], out[f"A_{tag}_dist"], out[f"A_{tag}_nr"] = i, d, n
    # case G: 24 chromosomes (gonosomal placeholder rows)
    per2 = list(per) + [20, 6]
    x2, per2, cum2 = synth.make_corrected_matrix(per2, 11, seed=2)
    random.seed(3)
    i, d, n = R.newref_tools.get_reference(x2, per2, cum2, 30, 1, 1)
    random.seed(3)
    out["G_ids"] = np.array(random.sample(range(11), 11))
    out["G_x"], out["G_per"], out["G_cum"] = x2, per2, cum2
    out["G_idx"], out["G_dist"], out["G_nr"] = i, d, n
    # case T: ties + fewer candidates than refsize + empty chromosomes
    per3 = np.array([6, 5, 4] + [0] * 19)
    x3 = np.round(synth.make_corrected_matrix(per3, 4, seed=5)[0] * 8) / 8
    cum3 = np.cumsum(per3)
    random.seed(1)
    i, d, n = R.newref_tools.get_reference(x3, per3, cum3, 12, 1, 1)
    random.seed(1)
    out["T_ids"] = np.array(random.sample(range(4), 4))
    out["T_x"], out["T_per"], out["T_cum"] = x3, per3, cum3
    out["T_idx"], out["T_dist"], out["T_nr"] = i, d, n
    # case S: S=129 (pairwise-sum recursion split) few rows
    per4 = [30, 25] + [3] * 20
    x4, per4, cum4 = synth.make_corrected_matrix(per4, 129, seed=9)
    random.seed(11)
    i, d, n = R.newref_tools.get_reference(x4, per4, cum4, 20, 1, 1)
    random.seed(11)
    out["S_ids"] = np.array(random.sample(range(129), 100))
    out["S_x"], out["S_per"], out["S_cum"] = x4, per4, cum4
    out["S_idx"], out["S_dist"], out["S_nr"] = i, d, n
    np.savez_compressed(os.path.join(HERE, "get_reference.npz"), **out)
    print("get_reference.npz written")


def golden_newref_predict(R):
    """Runs the reference's whole `newref` (main.tool_newref) on 24 synthetic samples at 5 Mb
    bins, then `normalize` (predict_control.py:21) for A / F / M on a test sample with a planted
    gain, and `get_z_score` (overall_tools.py:88) on hand-made segments."""
    ref_main = R.main  # ref_loader injects qc_reference (main.py:135 NameError)
    binsize = 5_000_000
    samples, genders = synth.make_samples(24, binsize, seed=21, depth=4e6)
    tmp = tempfile.mkdtemp()
    infiles = []
    for i, s in enumerate(samples):
        f = os.path.join(tmp, f"s{i}.npz")
        np.savez_compressed(f, binsize=binsize, sample=s, quality={})
        infiles.append(f)
    args = types.SimpleNamespace(infiles=infiles, outfile=os.path.join(tmp, "ref.npz"), nipt=False,
                                 yfrac=0.006, plotyfrac=None, refsize=60, binsize=binsize, cpus=1)
    random.seed(5)
    np.random.seed(5)
    try:
        ref_main.tool_newref(args)
    except SystemExit:
        pass
    ref = dict(np.load(args.outfile, allow_pickle=True))
    test_samples, _ = synth.make_samples(2, binsize, seed=99, depth=4e6,
                                         cnv=[(0, 3, 5, 25, 1.5), (1, 8, 0, 12, 0.5)])
    out = {("ref__" + k): v for k, v in ref.items()}
    pargs = types.SimpleNamespace(maskrepeats=5, minrefbins=10)
    for si, (ts, g) in enumerate(zip(test_samples, ["F", "M"])):
        ts = {k: v.copy() for k, v in ts.items()}
        for k, v in ts.items():
            out[f"t{si}_sample_{k}"] = v
        ts = R.overall_tools.gender_correct(ts, g)
        for rg in ["A", g]:
            r, z, w, n, m_lr, m_z = R.predict_control.normalize(pargs, ts, ref, rg)
            out[f"t{si}_{rg}_r"], out[f"t{si}_{rg}_z"], out[f"t{si}_{rg}_w"] = r, z, w
            out[f"t{si}_{rg}_n"] = n
            out[f"t{si}_{rg}_m"] = np.array([m_lr, m_z])
        out[f"t{si}_cutoff"] = np.array(R.predict_tools.get_optimal_cutoff(ref, 5))
    # get_z_score: results per chromosome over unmasked bins
    rng = np.random.default_rng(4)
    bpc = ref["bins_per_chr.F"]
    nr_m = ref["null_ratios"].shape[1]
    results_r, results_w, results_nr = [], [], []
    for c, nb in enumerate(bpc):
        r = rng.normal(0, 0.05, nb)
        r[rng.random(nb) < 0.15] = 0
        w = rng.uniform(0.5, 2.0, nb)
        nr = rng.normal(0, 0.05, (nb, nr_m))
        nr[rng.random((nb, nr_m)) < 0.02] = np.inf
        rows = [nr[i].tolist() if r[i] != 0 or rng.random() < 0.5 else 0 for i in range(nb)]
        results_r.append(r.tolist())
        results_w.append(w.tolist())
        results_nr.append(rows)
    segs = []
    for c, nb in enumerate(bpc):
        cuts = sorted(set([0, nb] + rng.integers(0, nb, 2).tolist()))
        for s, e in zip(cuts[:-1], cuts[1:]):
            segs.append([c, int(s), int(e), float(rng.normal(0, 0.05))])
    import copy
    zs = R.overall_tools.get_z_score(segs, {"results_nr": copy.deepcopy(results_nr),
                                            "results_r": results_r, "results_w": results_w})
    out["zs_segs"] = np.array(segs, dtype=float)
    out["zs_z"] = np.array([np.nan if isinstance(z, str) else z for z in zs])
    out["zs_r"] = np.concatenate([np.array(x) for x in results_r])
    out["zs_w"] = np.concatenate([np.array(x) for x in results_w])
    dense = np.full((len(out["zs_r"]), nr_m), np.nan)
    has = np.zeros(len(out["zs_r"]), dtype=bool)
    k = 0
    for rows in results_nr:
        for row in rows:
            if not isinstance(row, (int, float)):
                dense[k] = row
                has[k] = True
            k += 1
    out["zs_nr"], out["zs_has_nr"], out["zs_bpc"] = dense, has, np.array(bpc)
    np.savez_compressed(os.path.join(HERE, "newref_predict.npz"), **out)
    print("newref_predict.npz written", {k: v.shape for k, v in ref.items() if hasattr(v, "shape")})
    golden_tool_test(R, ref_main, args.outfile, test_samples, tmp, binsize)


def golden_tool_test(R, ref_main, ref_path, test_samples, tmp, binsize):
    """The reference's whole `tool_test` (main.py:145-300) on the two test samples against the reference .npz it
    built itself, with the R bridge replaced by tests/golden/stub_cbs.py (R / DNAcopy are absent):
    golden for the result assembly (main.py:242-271), get_post_processed_result, log_trans, apply_blacklist,
    get_z_score as called from exec_cbs, and the four output tables."""
    import copy
    from stub_cbs import stub_segments
    out = {}
    bl = os.path.join(tmp, "blacklist.bed")
    with open(bl, "w") as fh:
        fh.write("chr3\t{}\t{}\n".format(2 * binsize + 5, 4 * binsize - 1))
        fh.write("X\t{}\t{}\n".format(10 * binsize, 12 * binsize))
        fh.write("chr7\t0\t100\n")
    out["blacklist_text"] = np.array(open(bl).read())

    def fake_exec_R(json_dict):
        if "results_c" in json_dict:
            return None
        nchr = 24 if json_dict["ref_gender"] == "M" else 23  # CBS.R:30-34
        return [{"chr": c + 1, "s": s_, "e": e_, "r": r_} for c, s_, e_, r_ in
                stub_segments(json_dict["results_r"], json_dict["results_w"], nchr)]

    captured = {}
    orig_tables = ref_main.generate_output_tables

    def capture(rem_input, results):
        captured["rem_input"] = {k: v for k, v in rem_input.items() if k != "args"}
        captured["results"] = copy.deepcopy({k: v for k, v in results.items() if k != "results_nr"})
        orig_tables(rem_input, results)

    R.predict_tools.exec_R = fake_exec_R
    ref_main.generate_output_tables = capture
    try:
        for si, ts in enumerate(test_samples):
            f = os.path.join(tmp, f"test{si}.npz")
            np.savez_compressed(f, binsize=binsize, sample=ts, quality={})
            outid = os.path.join(tmp, f"out{si}")
            targs = types.SimpleNamespace(infile=f, reference=ref_path, outid=outid, minrefbins=10, maskrepeats=5, alpha=1e-4,
                                          zscore=5, beta=None, blacklist=bl if si == 0 else None, gender=None, ylim="def",
                                          bed=True, plot=False, cairo=False, add_plot_title=False, seed=1, regions=None)
            ref_main.tool_test(targs)
            res, rem = captured["results"], captured["rem_input"]
            for key in ("results_r", "results_z", "results_w"):
                out[f"t{si}_{key}"] = np.concatenate([np.asarray(x, dtype=float) for x in res[key]])
            out[f"t{si}_results_c"] = np.array([[c[0], c[1], c[2], np.nan if isinstance(c[3], str) else c[3], c[4]]
                                                for c in res["results_c"]], dtype=float)
            out[f"t{si}_meta"] = np.array([rem["ref_gender"], rem["gender"], str(rem["n_reads"]), str(rem["binsize"])])
            for sfx in ("_bins.bed", "_segments.bed", "_aberrations.bed", "_statistics.txt"):
                out[f"t{si}{sfx}"] = np.array(open(outid + sfx).read())
    finally:
        ref_main.generate_output_tables = orig_tables
    np.savez_compressed(os.path.join(HERE, "tool_test.npz"), **out)
    print("tool_test.npz written", sorted(out))


def golden_prep(R):
    """tool_newref_prep (newref_control.py:24-80) for the A, F and M passes of one sample set, the mask leaking from
    pass to pass as in main.py:98-137: normalize_and_mask, train_pca (np.random.seed pinned; sklearn's randomized
    solver where `auto` picks it), the PCA-distance filter with the in-place mask edit and the redo.
    Six bins carry sample-specific noise the five components cannot explain, so the filter fires."""
    binsize = 2_000_000
    S = 60
    samples, genders = synth.make_samples(S, binsize, seed=41, depth=8e6)
    rng = np.random.default_rng(42)
    lens = [len(samples[0][str(c)]) for c in range(1, 25)]
    offs = np.concatenate([[0], np.cumsum(lens)])
    for b in rng.choice(int(offs[22]), 6, replace=False):
        c = int(np.searchsorted(offs, b, side="right")) - 1
        for s in samples:
            s[str(c + 1)][b - offs[c]] = int(s[str(c + 1)][b - offs[c]] * rng.lognormal(0, 1.2))
    out = {"counts": np.stack([np.concatenate([s[str(c)] for c in range(1, 25)]) for s in samples], axis=1).astype(np.int32),
           "lens": np.array(lens), "genders": np.array(genders), "binsize": np.array(binsize)}
    samples = np.array(samples)
    for i, s in enumerate(samples):
        samples[i] = R.overall_tools.gender_correct(s, genders[i])
    total_mask, bpc = R.newref_tools.get_mask(samples)
    g = np.array(genders)
    total_mask = total_mask & R.newref_tools.get_mask(samples[g == "F"])[0] & R.newref_tools.get_mask(samples[g == "M"])[0]
    out["total_mask"], out["bins_per_chr"] = total_mask.copy(), np.array(bpc)
    tmp = tempfile.mkdtemp()
    for gender, sub in (("A", samples), ("F", samples[g == "F"]), ("M", samples[g == "M"])):
        a = types.SimpleNamespace(prepdatafile=os.path.join(tmp, "d.npy"), prepfile=os.path.join(tmp, "p.npz"), binsize=binsize)
        np.random.seed(3)
        R.newref_control.tool_newref_prep(a, sub, gender, total_mask, bpc)  # edits total_mask in place
        p = np.load(a.prepfile)
        for key in ("mask", "masked_bins_per_chr", "masked_bins_per_chr_cum", "pca_components", "pca_mean"):
            out[f"{gender}_{key}"] = p[key]
        out[f"{gender}_corrected_rows"] = np.load(a.prepdatafile)[::5]  # every 5th bin keeps the fixture small
        out[f"{gender}_total_mask_after"] = total_mask.copy()
    np.savez_compressed(os.path.join(HERE, "prep.npz"), **out)
    print("prep.npz written", {k: v.shape for k, v in out.items()})


def golden_example_bed():
    """Soft known answer for CBS: the per-bin log2 ratios of the reference's
    shipped example output docs/include/example.bed/ID_bins.bed (100 kb, T21 NIPT, older version)
    and its ID_segments.bed.  NaN -> 0 (the reference's "no data" sentinel)."""
    base = "/root/reference/docs/include/example.bed/"
    chrs, ratio = [], []
    for line in open(base + "ID_bins.bed").read().splitlines()[1:]:
        f = line.split("\t")
        c = {"X": 23, "Y": 24}.get(f[0], None) or int(f[0])
        chrs.append(c)
        ratio.append(0.0 if f[4] in ("NaN", "nan") else float(f[4]))
    segs = []
    for line in open(base + "ID_segments.bed").read().splitlines()[1:]:
        f = line.split("\t")
        c = {"X": 23, "Y": 24}.get(f[0], None) or int(f[0])
        segs.append([c, (int(f[1]) - 1) // 100000, int(f[2]) // 100000, float(f[3])])
    np.savez_compressed(os.path.join(HERE, "example_bed.npz"), chr=np.array(chrs, dtype=np.int8),
                        ratio=np.array(ratio, dtype=np.float32), segments=np.array(segs, dtype=np.float64))
    print("example_bed.npz written", len(ratio), len(segs))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    R = ref_loader.load()
    if a.only in (None, "get_reference"):
        golden_get_reference(R)
    if a.only in (None, "predict"):
        golden_newref_predict(R)
    if a.only in (None, "example_bed"):
        golden_example_bed()
    if a.only in (None, "prep"):
        golden_prep(R)
