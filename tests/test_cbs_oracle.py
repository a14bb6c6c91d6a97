"""CPU suite for the CBS restatement (oracle/cbs_oracle.py).  PARITY UNPINNED against DNAcopy (no R in
the container, no fixture in the reference): these tests pin the behaviour the published algorithm
guarantees and the one soft known answer the reference ships (docs/include/example.bed)."""
import os

import numpy as np
import pytest

from oracle import cbs_oracle as C


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    assert C.philox4x32((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert C.philox4x32((0xffffffff,) * 4, (0xffffffff, 0xffffffff)) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert C.philox4x32((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_tail_probability_monotone_and_scaled():
    p = [C.tailp(b, 26 / 5000.0, 5000) for b in (3.0, 4.0, 5.0, 6.0, 6.9)]
    assert all(a > b for a, b in zip(p[:-1], p[1:]))
    assert 1e-6 < C.tailp(6.0, 26 / 16000.0, 16000) < 1e-3
    assert abs(C.nu(0.005, 1e-6) - np.exp(-0.583 * 0.005)) < 1e-12
    assert 0.5 < C.nu(1.0, 1e-6) < 0.6  # Siegmund's nu(1) ~ 0.56


def test_planted_breakpoints_recovered_exactly():
    rng = np.random.default_rng(1)
    n = 600
    y = rng.normal(0, 0.05, n)
    y[200:260] += 0.25
    y[400:] -= 0.1
    w = rng.uniform(0.5, 2, n)
    assert C.segment_chromosome(y, w, nperm=200) == [200, 260, 400, 600]


def test_no_breakpoint_on_pure_noise():
    rng = np.random.default_rng(2)
    for n in (50, 180, 900):
        y = rng.normal(0, 0.05, n)
        w = rng.uniform(0.5, 2, n)
        assert C.segment_chromosome(y, w, alpha=1e-4, nperm=300, seed=3) == [n]


def test_flat_and_tiny_series():
    assert C.segment_chromosome(np.ones(40) * 0.3, np.ones(40)) == [40]
    assert C.segment_chromosome(np.array([0.1, 0.5, -0.2]), np.ones(3)) == [3]


def test_cbs_r_postprocessing_splits_long_na_runs_and_uses_weighted_means():
    # one chromosome, binsize 1 Mb -> NA runs longer than int(2e6/1e6) = 2 split a segment (CBS.R:95)
    rng = np.random.default_rng(5)
    r = rng.normal(0.0, 0.01, 80)
    r[r == 0] = 1e-3
    r[30:36] = 0  # run of 6 NA inside a flat segment -> split
    r[50:52] = 0  # run of 2 NA -> no split (not > 2)
    w = rng.uniform(0.5, 2.0, 80)
    w[10] = 0  # weight 0 -> 1 (CBS.R:42)
    rr = [r.tolist()] + [[0.0] * 5 for _ in range(22)]
    ww = [w.tolist()] + [[1.0] * 5 for _ in range(22)]
    out = C.cbs_r(rr, ww, "F", 1e-4, 1e6, nperm=100)
    assert [(d["chr"], d["s"], d["e"]) for d in out] == [(1, 0, 30), (1, 35, 80)]
    # right piece starts on the LAST NA bin of the run; its mean ignores NA bins
    w2 = w.copy(); w2[10] = 1.0
    m = r[35:80] != 0
    np.testing.assert_allclose(out[1]["r"], np.sum(r[35:80][m] * w2[35:80][m]) / np.sum(w2[35:80][m]), rtol=1e-14)
    m0 = r[0:30] != 0
    np.testing.assert_allclose(out[0]["r"], np.sum(r[0:30][m0] * w2[0:30][m0]) / np.sum(w2[0:30][m0]), rtol=1e-14)


def test_example_bed_chr21_gain(golden_dir):
    """The reference's shipped example (T21): an unweighted CBS of the published per-bin ratios
    finds the chr21 gain 13100001-46700000 with ratio 0.0923 (docs/include/example.bed)."""
    g = np.load(os.path.join(golden_dir, "example_bed.npz"))
    chrs, ratio = g["chr"].astype(int), g["ratio"].astype(np.float64)
    rr = [ratio[chrs == c].tolist() for c in range(1, 24)]
    ww = [np.ones(int((chrs == c).sum())).tolist() for c in range(1, 24)]
    out = C.cbs_r(rr, ww, "F", 1e-4, 100000, nperm=200)
    c21 = [d for d in out if d["chr"] == 21]
    want = [s for s in g["segments"] if int(s[0]) == 21]
    gain = max(c21, key=lambda d: d["e"] - d["s"])
    ref_gain = max(want, key=lambda s: s[2] - s[1])
    assert (gain["s"], gain["e"]) == (int(ref_gain[1]), int(ref_gain[2]))
    assert abs(gain["r"] - ref_gain[3]) < 2e-3  # ratios in the bed file are rounded to 4 decimals
    # genome-wide: all 50 segments of the published result (docs/include/example.bed/ID_segments.bed) are reproduced,
    # start and end, from the published per-bin ratios
    got = {(d["chr"], d["s"], d["e"]) for d in out}
    ref = {(int(s[0]), int(s[1]), int(s[2])) for s in g["segments"] if int(s[0]) <= 23}
    assert got == ref, (sorted(ref - got), sorted(got - ref))


def test_sequential_boundary_table():
    """getbdry restated (oracle) == the product's table (wisecondorx_b200.cbs.sequential_boundary); structure of the
    table: first entry nperm - int(nperm * eta), rows ascending, the early-stop probability of every row about eta."""
    from wisecondorx_b200 import cbs
    for eta, nperm, ones in [(0.05, 10000, 2), (0.05, 10000, 5), (0.05, 400, 9), (0.1, 1000, 3)]:
        t = C.seq_boundary(eta, nperm, ones)
        assert t == cbs.sequential_boundary(eta, nperm, ones).tolist()
        assert len(t) == ones * (ones + 1) // 2 and t[0] == nperm - int(nperm * eta)
        for j in range(1, ones + 1):
            row = t[j * (j - 1) // 2: j * (j + 1) // 2]
            assert row == sorted(row) and 1 <= row[0] and row[-1] <= nperm
            assert abs(C._p_exceed(nperm, j, row) - eta) < 0.02
    # closed form for one tolerated exceedance (two ones): C(N - b1, 2) + b1 (N - b2) placements stop early
    from math import comb
    b1, b2 = C.seq_boundary(0.05, 10000, 2)[1:]
    assert abs((comb(10000 - b1, 2) + b1 * (10000 - b2)) / comb(10000, 2) - C._p_exceed(10000, 2, [b1, b2])) < 1e-9


def test_product_postprocessing_equals_oracle(monkeypatch):
    """wisecondorx_b200.cbs.cbs_segments_batch -- CBS.R:30-129 around the device call, on host threads in the library
    (csrc/host_cbs.cu: runs here, it needs no GPU) -- against the oracle's cbs_r (per segment, like CBS.R) on the same
    random segment ends: NA runs at chromosome ends, runs longer and shorter than the split threshold, all-NA
    chromosomes, weight 0, every bin size class, single samples and a batch."""
    from wisecondorx_b200 import cbs
    rng = np.random.default_rng(5)
    total = 0
    cases = []
    for trial in range(24):
        binsize = [15000.0, 100000.0, 5000.0, 1e6][trial % 4]
        g = "M" if trial % 5 == 0 else "F"
        per = [int(x) for x in rng.integers(5, 1500, 24 if g == "M" else 23)]
        rr = [rng.normal(0, 0.1, n) for n in per]
        ww = [rng.uniform(0.5, 2, n) for n in per]
        for r in rr:
            r[rng.random(len(r)) < [0.0, 0.05, 0.3][trial % 3]] = 0
            for _ in range(trial % 5):
                a = int(rng.integers(0, max(1, len(r) - 1)))
                r[a:a + int(rng.integers(1, 400))] = 0
        rr[0][:3] = 0
        rr[0][-2:] = 0
        rr[7][:] = 0  # CBS.R:56-63: dropped
        ww[3][::5] = 0  # CBS.R:42
        cases.append((rr, ww, g, binsize))

    def run(batch, binsize):
        """Product and oracle around the same segmenter: ends drawn per (sample, chromosome), remembered by series content."""
        ends_of = {}

        def key(yy, c):
            return (int(c), len(yy), float(np.sum(yy)))

        def fake_segment_flat(y, w, off, ids, alpha, nperm, seed, ctx, sequential=True, eta=0.05):
            ends, nseg = [], []
            for s in range(len(off) - 1):
                yy = y[off[s]:off[s + 1]]
                n = len(yy)
                e = sorted(set([int(x) for x in rng.integers(1, n + 1, int(rng.integers(0, 6)))] + [n]))
                ends_of[key(yy, ids[s])] = e
                ends += e
                nseg.append(len(e))
            return np.array(ends, dtype=np.int32), np.array(nseg, dtype=np.int32)

        monkeypatch.setattr(cbs, "_segment_flat", fake_segment_flat)
        got = cbs.cbs_segments_batch([(rr, ww, g) for rr, ww, g, _ in batch], 1e-4, binsize, seed=1)
        n_out = 0
        for (rr, ww, g, _), segs in zip(batch, got):
            want = [[d["chr"] - 1, d["s"], d["e"], d["r"]] for d in
                    C.cbs_r(rr, ww, g, 1e-4, binsize, segmenter=lambda yy, wv, c: ends_of[key(yy, c)])]
            assert len(segs) == len(want)
            for a, b in zip(segs, want):
                assert a[:3] == b[:3] and (a[3] == b[3] or (np.isnan(a[3]) and np.isnan(b[3]))), (a, b)
                assert type(a[0]) is int and type(a[1]) is int and type(a[2]) is int and type(a[3]) is float
            n_out += len(segs)
        return n_out

    for case in cases:
        total += run([case], case[3])
    for bs in (15000.0, 100000.0, 5000.0, 1e6):
        total += run([c for c in cases if c[3] == bs], bs)
    assert total > 2000
