"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper around oracle/wcx_oracle.c (the plain-C
restatement of get_reference).  Built by ``oracle/Makefile`` / ``build()`` with
``gcc -O2 -ffp-contract=off -fopenmp``; the .so is git-ignored but travels to the GPU box."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libwcx_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "wcx_oracle.c")
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC",
                               src, "-o", SO, "-lm"])
    return SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            build()
        L = ctypes.CDLL(SO)
        vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
        L.wcxo_topk.restype = ctypes.c_int
        L.wcxo_topk.argtypes = [vp, i64, i32, vp, vp, i32, i64, i64, i32, vp, vp, i32]
        L.wcxo_null_ratios.restype = ctypes.c_int
        L.wcxo_null_ratios.argtypes = [vp, i64, i32, vp, i64, i64, i32, vp, i32, vp, i32]
        L.wcxo_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def max_threads() -> int:
    return int(lib().wcxo_max_threads())


def topk(x, per, cum, k, row_begin, row_end, nthreads=0):
    x = np.ascontiguousarray(x, dtype=np.float64)
    per = np.ascontiguousarray(per, dtype=np.int64)
    cum = np.ascontiguousarray(cum, dtype=np.int64)
    rows = row_end - row_begin
    idx = np.empty((rows, k), dtype=np.int32)
    dist = np.empty((rows, k), dtype=np.float64)
    nt = nthreads or max_threads()
    rc = lib().wcxo_topk(x.ctypes.data, x.shape[0], x.shape[1], per.ctypes.data, cum.ctypes.data,
                         len(cum), row_begin, row_end, k, idx.ctypes.data, dist.ctypes.data, nt)
    if rc:
        raise RuntimeError("wcxo_topk failed")
    return idx, dist


def null_ratios(x, idx, row_begin, row_end, sample_ids, nthreads=0):
    x = np.ascontiguousarray(x, dtype=np.float64)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    ids = np.ascontiguousarray(sample_ids, dtype=np.int32)
    out = np.empty((row_end - row_begin, len(ids)), dtype=np.float64)
    nt = nthreads or max_threads()
    lib().wcxo_null_ratios(x.ctypes.data, x.shape[0], x.shape[1], idx.ctypes.data, row_begin,
                           row_end, idx.shape[1], ids.ctypes.data, len(ids), out.ctypes.data, nt)
    return out


def get_reference(x, per, cum, ref_size, part, split_parts, sample_ids, nthreads=0):
    n = int(cum[-1])
    start = int(n / float(split_parts) * (part - 1))
    end = int(n / float(split_parts) * part)
    idx, dist = topk(x, per, cum, ref_size, start, end, nthreads)
    nr = null_ratios(x, idx, start, end, sample_ids, nthreads)
    return idx, dist, nr
