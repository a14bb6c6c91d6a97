"""CPU suite: the C-ABI library loads and exports every symbol include/wcx_b200.h declares; host
logic of the reference-facing mirror (no compute calls: there is no GPU here)."""
import ctypes
import os

import numpy as np
import pytest

from wisecondorx_b200 import _lib, newref_tools


@pytest.fixture(scope="module")
def lib():
    from wisecondorx_b200 import build
    build.build()
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    names = _lib.declared_symbols()
    assert "wcx_newref_topk" in names and "wcx_get_reference" in names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/wcx_b200.h but not exported"


def test_version_and_error_string(lib):
    assert lib.wcx_version() >= 100
    assert isinstance(lib.wcx_last_error(), bytes)


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.WcxError):
        _lib.Context(0)


@pytest.mark.parametrize("n,parts", [(2815, 1), (2815, 8), (27941, 7), (191678, 8)])
def test_get_part_matches_reference_arithmetic(n, parts):
    # reference newref_tools.py:244-247: int(bincount / float(outof) * partnum)
    covered = 0
    for p in range(parts):
        s, e = newref_tools._get_part(p, parts, n)
        assert s == int(n / float(parts) * p) and e == int(n / float(parts) * (p + 1))
        assert s == covered
        covered = e
    assert covered == n


def test_host_entry_points_reject_bad_arguments(lib):
    """The host-side entry points follow the library's error convention (non-zero + wcx_last_error) instead of reading
    through null pointers or past short buffers."""
    p = lambda a: ctypes.c_void_p(a.ctypes.data)  # noqa: E731
    one = np.zeros(4)
    n64 = ctypes.c_int64()
    assert lib.wcx_host_bin_sums(None, 3, 2, None, 0, p(one), p(one), 1) != 0 and b"wcx_host_bin_sums" in lib.wcx_last_error()
    counts = np.ones((3, 2), dtype=np.int32)
    bad_cols = np.array([0, 2], dtype=np.int32)
    assert lib.wcx_host_bin_sums(p(counts), 3, 2, p(bad_cols), 2, p(one), p(one), 1) != 0 and b"out of range" in lib.wcx_last_error()
    assert lib.wcx_host_stack_counts(None, None, 2, 3, None, None, 1) != 0
    # a sample longer than the rows of its chromosome
    col = np.arange(5, dtype=np.int32)
    ptrs = np.array([col.ctypes.data], dtype=np.uintp)
    lens, offs = np.array([5], dtype=np.int64), np.array([0, 4], dtype=np.int64)
    out = np.zeros((4, 1), dtype=np.int32)
    assert lib.wcx_host_stack_counts(p(ptrs), p(lens), 1, 1, p(offs), p(out), 1) != 0 and b"longer" in lib.wcx_last_error()
    assert lib.wcx_predict_assemble(None, None, None, 4, None, None, None, None, 0, None, None, None, 1, 150.0, None, 4,
                                    None, None, None, None, 1) != 0
    # fewer results than kept bins: return code 2 (the reference's IndexError)
    mask = np.ones(6, dtype=np.uint8)
    rows = np.zeros(1, dtype=np.int32)
    o_r, o_z, o_w, o_i = np.zeros(6), np.zeros(6), np.zeros(6), np.zeros(6, dtype=np.int32)
    assert lib.wcx_predict_assemble(p(one), p(one), p(one), 4, p(rows), p(one), p(one), p(one), 1, p(np.ones(5)), p(one), p(one),
                                    1, 150.0, p(mask), 6, p(o_r), p(o_z), p(o_w), p(o_i), 1) == 2
    assert lib.wcx_cbs_pack_count(None, None, None, 1, None, 1) != 0 and lib.wcx_cbs_unpack(*([None] * 4), 1, *([None] * 4), 3, *([None] * 4), 1) != 0
    small = np.zeros(8, dtype=np.uint8)
    assert lib.wcx_host_format_repr(p(one), 4, p(small), small.nbytes, ctypes.byref(n64)) != 0
    assert lib.wcx_host_format_bins(b"1", 1000, p(one), p(one), 4, p(small), small.nbytes, ctypes.byref(n64)) != 0
    # segment ends out of order
    pos = np.arange(6, dtype=np.int32)
    y = np.ones(6)
    off = np.array([0, 6], dtype=np.int64)
    ends, nseg = np.array([4, 3], dtype=np.int32), np.array([2], dtype=np.int32)
    slot, start = np.array([0, 4], dtype=np.int64), np.zeros(1, dtype=np.int64)
    os_, oa, ob, orr = np.zeros(4, dtype=np.int32), np.zeros(4, dtype=np.int64), np.zeros(4, dtype=np.int64), np.zeros(4)
    assert lib.wcx_cbs_unpack(p(pos), p(y), p(y), p(off), 1, p(ends), p(nseg), p(start), p(slot), 10, p(os_), p(oa), p(ob), p(orr), 1) != 0
    assert b"out of order" in lib.wcx_last_error()
