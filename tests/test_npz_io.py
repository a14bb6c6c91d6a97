"""The parallel .npz writer / loader (wisecondorx_b200/npz_io.py) against NumPy's own reader and writer:
same keys, dtypes, shapes and values; a valid zip archive; multi-block deflate streams; chained CRC-32."""
import zipfile
import zlib

import numpy as np
import pytest

from wisecondorx_b200 import npz_io


def test_crc32_combine_matches_zlib():
    rng = np.random.default_rng(0)
    for la, lb in [(0, 5), (7, 0), (1, 1), (1000, 12345), (65536, 3)]:
        a, b = rng.bytes(la), rng.bytes(lb)
        assert npz_io._crc32_combine(zlib.crc32(a), zlib.crc32(b), lb) == zlib.crc32(a + b)


def test_savez_compressed_roundtrip(tmp_path, monkeypatch):
    monkeypatch.setattr(npz_io, "BLOCK", 1 << 16)  # force many blocks per member
    rng = np.random.default_rng(1)
    d = {"indexes": rng.integers(0, 1000, (3000, 300), dtype=np.int32), "distances": rng.random((3000, 300)),
         "null_ratios.F": rng.standard_normal((3000, 100)), "mask": rng.random(5000) > 0.03, "binsize": 15000,
         "has_female": True, "trained_cutoff": np.float64(0.0023), "empty": np.zeros((0, 3)),
         "sample": np.array({"1": np.arange(5, dtype=np.int32)}, dtype=object)}
    ours, ref = tmp_path / "ours.npz", tmp_path / "ref.npz"
    npz_io.savez_compressed(str(ours), threads=4, **d)
    np.savez_compressed(str(ref), **d)
    assert zipfile.ZipFile(ours).testzip() is None
    a, b = np.load(ours, allow_pickle=True), np.load(ref, allow_pickle=True)
    assert sorted(a.files) == sorted(b.files)
    for src in (ours, ref):  # the concurrent reader on both writers' files
        c = npz_io.load_npz(str(src), threads=3)
        assert sorted(c) == sorted(b.files)
        for k in b.files:
            assert c[k].dtype == b[k].dtype and c[k].shape == b[k].shape, k
            if c[k].dtype != object:
                assert np.array_equal(c[k], b[k]), k
    for k in b.files:
        assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape, k
        if a[k].dtype == object:
            assert np.array_equal(a[k].item()["1"], b[k].item()["1"])
        else:
            assert np.array_equal(a[k], b[k]), k


def test_load_samples(tmp_path):
    paths = []
    for i in range(5):
        p = tmp_path / f"s{i}.npz"
        np.savez_compressed(str(p), binsize=5000, sample={"1": np.full(4, i, dtype=np.int32)}, quality={})
        paths.append(str(p))
    out = npz_io.load_samples(paths, threads=3)
    assert [int(s["1"][0]) for s, _ in out] == list(range(5)) and all(b == 5000 for _, b in out)


def test_post_processed_result_matches_reference_inflate_loop():
    """main.get_post_processed_result against a restatement of the reference's loops (predict_control.py:49-63,
    predict_tools.py:163-170), including more results than kept bins (surplus ignored)."""
    from wisecondorx_b200 import main as wmain
    rng = np.random.default_rng(5)
    bins_per_chr = [7, 5, 4]
    mask = rng.random(16) > 0.3
    cnt = int(mask.sum())
    for extra in (0, 3):
        res = rng.random(cnt + extra)
        sizes = rng.integers(0, 300, cnt + extra)
        want_flat = [0.0] * len(mask)
        r = res.copy()
        r[sizes < 150] = 0
        j = 0
        for i, v in enumerate(mask):
            if v:
                want_flat[i] = r[j]
                j += 1
        want = [want_flat[sum(bins_per_chr[:c]):sum(bins_per_chr[:c + 1])] for c in range(3)]
        got = wmain.get_post_processed_result(150, res, sizes, mask, bins_per_chr)
        assert all(np.array_equal(g, np.array(w)) for g, w in zip(got, want))


def test_async_writer_and_newref_merge(tmp_path):
    """AsyncNpzWriter (arrays deflated while the caller goes on) through tool_newref_merge: key layout of the
    reference's final .npz (newref_control.py:220-237), passes queued early or at merge time."""
    from wisecondorx_b200 import newref_control
    rng = np.random.default_rng(3)

    def fake_pass(gender, n):
        return {"gender": gender, "mask": rng.random(50) > 0.1, "bins_per_chr": np.arange(3), "masked_bins_per_chr": np.arange(3),
                "masked_bins_per_chr_cum": np.cumsum(np.arange(3)), "pca_components": rng.random((5, n)), "pca_mean": rng.random(n),
                "indexes": rng.integers(0, n, (n, 7), dtype=np.int32), "distances": rng.random((n, 7)), "null_ratios": rng.random((n, 4))}

    for early in (False, True):
        out = str(tmp_path / f"ref{int(early)}.npz")
        results = [fake_pass("A", 40), fake_pass("F", 45), fake_pass("M", 47)]
        writer = npz_io.AsyncNpzWriter(out) if early else None
        if early:
            for r in results[:2]:
                newref_control.writer_add_pass(writer, r, 15000)
        newref_control.tool_newref_merge(out, results, 15000, False, 0.0023, writer)
        z = np.load(out, allow_pickle=True)
        for res, sfx in zip(results, ("", ".F", ".M")):
            assert int(z["binsize" + sfx]) == 15000
            for key in newref_control.RESULT_KEYS:
                assert np.array_equal(z[key + sfx], res[key]) and z[key + sfx].dtype == np.asarray(res[key]).dtype
        assert bool(z["has_female"]) and bool(z["has_male"]) and not bool(z["is_nipt"]) and float(z["trained_cutoff"]) == 0.0023
        assert "_queued" not in z.files and "gender" not in z.files
        assert len(z.files) == 3 * 10 + 4


def test_load_samples_fast_reader_equals_numpy(tmp_path):
    """Sample files are parsed without the zipfile module (npz_io._read_members); same dicts as np.load, compressed or
    stored, CRC checked."""
    from wisecondorx_b200 import synth
    samples, _ = synth.make_samples(6, 1000000, seed=5)
    paths = []
    for i, smp in enumerate(samples):
        pth = str(tmp_path / ("s%d.npz" % i))
        (np.savez_compressed if i % 2 == 0 else np.savez)(pth, binsize=1000000, sample=smp, quality={"x": i})
        paths.append(pth)
    assert npz_io._read_members(paths[0], ("sample", "binsize")) is not None
    for pth, (smp, bs) in zip(paths, npz_io.load_samples(paths)):
        with np.load(pth, encoding="latin1", allow_pickle=True) as z:
            want = z["sample"].item()
            assert bs == int(z["binsize"])
        assert smp.keys() == want.keys()
        for key in want:
            assert np.array_equal(smp[key], want[key]) and smp[key].dtype == want[key].dtype
    raw = bytearray(open(paths[0], "rb").read())
    raw[80] ^= 0xFF  # inside the first member's deflate stream
    bad = str(tmp_path / "bad.npz")
    open(bad, "wb").write(bytes(raw))
    with pytest.raises(Exception):
        npz_io.load_samples([bad])
