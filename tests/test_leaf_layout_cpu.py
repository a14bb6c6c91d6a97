"""CPU test of the parity-critical host logic of the exact re-rank: the summation plan and the leaf-major row layout
built by the library (rerank.cu build_sum_plan / build_leaf_layout, exported host-only through
wcx_debug_leaf_layout) must reproduce np.sum(np.power(a - b, 2)) bit for bit when evaluated the way the kernel
evaluates them (8 accumulator chains per leaf, pairwise combination, sequential tail, leaves combined by the plan).
The GPU kernel is compared with the oracle in tests/test_newref_gpu.py; this pins the layout without a GPU."""
import ctypes

import numpy as np
import pytest

from wisecondorx_b200 import _lib


def _layout(s):
    from wisecondorx_b200 import build
    build.build()  # no-op when the in-tree library is up to date (nvcc cross-compiles without a GPU)
    L = _lib.load()
    sizes = np.zeros(3, dtype=np.int32)
    _lib.check(L.wcx_debug_leaf_layout(s, None, 0, None, 0, None, 0, ctypes.c_void_p(sizes.ctypes.data)))
    sp, nl, pl = (int(v) for v in sizes)
    perm = np.zeros(sp, dtype=np.int32); desc = np.zeros(4 * nl, dtype=np.int32); plan = np.zeros(3 * pl, dtype=np.int32)
    _lib.check(L.wcx_debug_leaf_layout(s, ctypes.c_void_p(perm.ctypes.data), sp, ctypes.c_void_p(desc.ctypes.data), 4 * nl,
                                       ctypes.c_void_p(plan.ctypes.data), 3 * pl, ctypes.c_void_p(sizes.ctypes.data)))
    return perm, desc.reshape(-1, 4), plan.reshape(-1, 3)


def _kernel_order_distance(a, b, perm, desc, plan):
    ap = np.where(perm >= 0, a[np.maximum(perm, 0)], 0.0)
    bp = np.where(perm >= 0, b[np.maximum(perm, 0)], 0.0)
    leaf_sums = []
    for off, steps, tail, _ in desc:
        r = np.zeros(8)
        for t in range(steps):
            for c in range(8):
                for e in range(2):
                    u = bp[off + t * 16 + c * 2 + e] - ap[off + t * 16 + c * 2 + e]
                    r[c] = r[c] + u * u
        s1 = [r[c] + r[c ^ 1] for c in range(8)]
        s2 = [s1[c] + s1[c ^ 2] for c in range(8)]
        res = s2[0] + s2[4]
        for i in range(tail):
            u = bp[off + steps * 16 + i] - ap[off + steps * 16 + i]
            res = res + u * u
        leaf_sums.append(res)
    stack, li = [], 0
    for op, _, _ in plan:
        if op == 0:
            stack.append(leaf_sums[li]); li += 1
        else:
            r_ = stack.pop(); l_ = stack.pop(); stack.append(l_ + r_)
    assert len(stack) == 1 and li == len(leaf_sums)
    return stack[0]


@pytest.mark.parametrize("s", [1, 5, 8, 9, 20, 37, 100, 128, 129, 136, 257, 500, 777, 1000])
def test_leaf_layout_reproduces_numpy_sum(s):
    perm, desc, plan = _layout(s)
    assert sorted(int(q) for q in perm if q >= 0) == list(range(s))          # a permutation of the columns + padding
    assert len(perm) % 16 == 0 and all(int(d[0]) % 16 == 0 for d in desc)    # every leaf starts on a 128-byte line
    assert all(0 <= int(d[1]) <= 8 and 0 <= int(d[2]) < 8 for d in desc)
    rng = np.random.default_rng(s)
    for _ in range(10):
        a = 1.0 + 0.05 * rng.standard_normal(s)
        b = 1.0 + 0.05 * rng.standard_normal(s)
        want = np.sum(np.power(np.array([b]) - a, 2), 1)[0]                   # newref_tools.py:260
        assert _kernel_order_distance(a, b, perm, desc, plan) == want
