"""wisecondorx_b200 -- B200-native numeric core of WisecondorX behind the reference's own
function signatures and .npz formats.  Host code is Python; all arithmetic runs in hand-written
sm_100a CUDA kernels reached through the C-ABI in include/wcx_b200.h (ctypes, _lib.py)."""
__version__ = "0.1.0"
