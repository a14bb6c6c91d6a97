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
