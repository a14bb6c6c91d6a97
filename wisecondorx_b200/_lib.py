"""ctypes binding of libwcx_b200.so (include/wcx_b200.h).  There is no CPU fallback: if the
library is missing or no sm_100 device is usable, every entry point raises."""
from __future__ import annotations

import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libwcx_b200.so")
HEADER = os.path.abspath(os.path.join(HERE, "..", "include", "wcx_b200.h"))

KERNEL_AUTO, KERNEL_TC, KERNEL_SIMT, KERNEL_EXACT, KERNEL_TC2, KERNEL_TC2H, KERNEL_TCH = 0, 1, 2, 3, 4, 5, 6

_lib = None


class WcxError(RuntimeError):
    pass


def declared_symbols(header: str = HEADER):
    """Function names declared in include/wcx_b200.h (used by the symbol-export test)."""
    txt = open(header).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(wcx_[a-z0-9_]+)\s*\(", txt)))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise WcxError(f"{SO_PATH} not built: run `python -m wisecondorx_b200.build` "
                       "(or __graft_entry__.build()); there is no CPU fallback")
    L = ctypes.CDLL(SO_PATH)
    vp, i64, i32, dp = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p
    L.wcx_version.restype = ctypes.c_int
    L.wcx_last_error.restype = ctypes.c_char_p
    L.wcx_create.argtypes = [i32, ctypes.POINTER(vp)]
    L.wcx_destroy.argtypes = [vp]
    L.wcx_destroy.restype = None
    L.wcx_set_stream.argtypes = [vp, vp]
    L.wcx_sync.argtypes = [vp]
    L.wcx_host_alloc.argtypes = [ctypes.c_uint64, ctypes.POINTER(vp)]
    L.wcx_host_free.argtypes = [vp]
    L.wcx_host_stack_counts.argtypes = [vp, vp, i32, i32, vp, vp, i32]
    L.wcx_host_bin_sums.argtypes = [vp, i64, i32, vp, i32, vp, vp, i32]
    L.wcx_host_format_bins.argtypes = [ctypes.c_char_p, i64, vp, vp, i64, vp, i64, ctypes.POINTER(i64)]
    L.wcx_host_format_repr.argtypes = [vp, i64, vp, i64, ctypes.POINTER(i64)]
    L.wcx_newref_load.argtypes = [vp, dp, i64, i32, vp, vp, i32, i32]
    L.wcx_newref_topk.argtypes = [vp, i64, i64, i32, i32, vp, vp, i32]
    L.wcx_newref_null_ratios.argtypes = [vp, vp, i32, i64, i64, i32, vp, i32, vp, i32]
    L.wcx_newref_reference.argtypes = [vp, i64, i64, i32, i32, vp, i32, vp, vp, vp, i32]
    L.wcx_get_reference.argtypes = [vp, dp, i64, i32, vp, vp, i32, i32, i64, i64, vp, i32, i32, vp, vp, vp]
    L.wcx_newref_stats.argtypes = [vp, vp]
    L.wcx_newref_stage_ms.argtypes = [vp, vp]
    L.wcx_debug_prep.argtypes = [vp, vp, vp, vp]
    L.wcx_debug_tc_tile_f16.argtypes = [vp, i64, i64, vp]
    L.wcx_debug_prep_f16.argtypes = [vp, vp, vp, vp, vp]
    L.wcx_debug_list_counts.argtypes = [vp, vp, i64]
    f64 = ctypes.c_double
    L.wcx_predict_load_ref.argtypes = [vp, i32, vp, vp, i64, i32, vp, vp, i32, vp, vp, i32, vp, i64]
    L.wcx_predict_weights.argtypes = [vp, i32, vp]
    L.wcx_predict_optimal_cutoff.argtypes = [vp, i32, i32, ctypes.POINTER(f64)]
    L.wcx_predict_normalize.argtypes = [vp, i32, vp, i32, f64, i32, i64, vp, vp, vp, vp, vp]
    L.wcx_segment_zscore.argtypes = [vp, vp, i64, i32, vp, vp, vp, i64, vp, vp, i32, vp]
    L.wcx_predict_stage_ms.argtypes = [vp, vp]
    L.wcx_predict_assemble.argtypes = [vp, vp, vp, i64, vp, vp, vp, vp, i64, vp, vp, vp, i32, f64, vp, i64, vp, vp, vp, vp, i32]
    L.wcx_cbs_segment.argtypes = [vp, vp, vp, vp, i32, vp, f64, i32, ctypes.c_uint32, vp, vp]
    L.wcx_cbs_stats.argtypes = [vp, vp]
    L.wcx_cbs_set_boundary.argtypes = [vp, vp, i32]
    L.wcx_cbs_pack_count.argtypes = [vp, vp, vp, i32, vp, i32]
    L.wcx_cbs_pack.argtypes = [vp, vp, vp, vp, i32, vp, i64, vp, vp, vp, vp, i32]
    L.wcx_cbs_unpack.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, vp, i64, vp, vp, vp, vp, i32]
    L.wcx_newref_normalize_and_mask.argtypes = [vp, vp, i64, i32, vp, i64, vp, i32]
    L.wcx_pca_gram.argtypes = [vp, vp, i64, i32, i32, vp, vp]
    L.wcx_prep_fetch.argtypes = [vp, i32, i64, i32, vp]
    L.wcx_prep_device_ptr.argtypes = [vp, i32, i64, i32, ctypes.POINTER(vp)]
    L.wcx_debug_leaf_layout.argtypes = [i32, vp, i32, vp, i32, vp, i32, vp]
    L.wcx_pca_apply.argtypes = [vp, vp, vp, i32, vp, vp, i32]
    L.wcx_pca_distance.argtypes = [vp, vp, i64, i32, i32, vp, vp]
    L.wcx_newref_prep_stage_ms.argtypes = [vp, vp]
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise WcxError(load().wcx_last_error().decode("utf-8", "replace"))


class Context:
    """One wcx_ctx (one GPU, one stream)."""

    def __init__(self, device: int = 0):
        L = load()
        h = ctypes.c_void_p()
        check(L.wcx_create(device, ctypes.byref(h)))
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            load().wcx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        if not self._h:
            raise WcxError("context closed")
        return self._h

    def set_stream(self, cuda_stream_ptr: int):
        check(load().wcx_set_stream(self.handle, ctypes.c_void_p(cuda_stream_ptr)))

    def sync(self):
        check(load().wcx_sync(self.handle))


_default_ctx = {}
_ctx_lock = __import__("threading").RLock()


def default_context(device: int = 0) -> Context:
    with _ctx_lock:
        if device not in _default_ctx:
            _default_ctx[device] = Context(device)
        return _default_ctx[device]


class PinnedPool:
    """NumPy arrays over page-locked host memory (wcx_host_alloc).  A buffer goes back to the pool when the last
    array that views it is garbage collected, so steady-state callers (a loop of predict batches) allocate nothing;
    cudaHostAlloc itself costs ~0.1-0.3 ms per MB, more than the copy it speeds up."""

    GRAIN = 1 << 20
    SLACK = 1.25  # a free buffer up to this factor larger than the request is reused (reserve() sizes are upper bounds)

    def __init__(self):
        self._free = {}

    def _size(self, nbytes):
        return max(self.GRAIN, (nbytes + self.GRAIN - 1) // self.GRAIN * self.GRAIN)

    def _alloc(self, size):
        p = ctypes.c_void_p()
        check(load().wcx_host_alloc(size, ctypes.byref(p)))
        return p.value

    def reserve(self, nbytes):
        """Page-locks a buffer of at least nbytes ahead of its use (e.g. from a background thread while the host is
        busy with something else) and leaves it in the pool."""
        size = self._size(int(nbytes))
        self._free.setdefault(size, []).append(self._alloc(size))

    def empty(self, shape, dtype=None):
        import weakref
        import numpy as np
        dtype = np.dtype(dtype or np.float64)
        nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        size = self._size(nbytes)
        ptr = None
        for have in sorted(self._free):
            if have >= size and have <= size * self.SLACK and self._free[have]:
                try:
                    ptr, size = self._free[have].pop(), have
                    break
                except IndexError:  # another thread took it
                    continue
        if ptr is None:
            ptr = self._alloc(size)
        lst = self._free.setdefault(size, [])
        buf = (ctypes.c_char * size).from_address(ptr)
        weakref.finalize(buf, lst.append, ptr)
        return np.frombuffer(buf, dtype=dtype, count=nbytes // dtype.itemsize).reshape(shape)

    def trim(self):
        """Frees the buffers that are not in use."""
        L = load()
        for lst in self._free.values():
            while lst:
                L.wcx_host_free(ctypes.c_void_p(lst.pop()))


pinned = PinnedPool()


def prewarm_async(device: int = 0, pinned_bytes=()):
    """Creates the CUDA context of `device` and page-locks result buffers on a background thread, so that a command
    line that starts with seconds of host-only work (reading samples, masks) does not pay for them on its critical
    path.  Returns the thread (join() is optional: default_context / pinned.empty are safe to call concurrently --
    the context is created once under a lock, a buffer that is not there yet is simply allocated by the caller)."""
    import threading

    def work():
        try:
            with _ctx_lock:
                default_context(device)
            for nb in pinned_bytes:
                pinned.reserve(nb)
        except Exception:  # the foreground call will raise the real error
            pass

    t = threading.Thread(target=work, name="wcx-prewarm", daemon=True)
    t.start()
    return t
