"""Host-side profile of `predict_control.predict_batch` WITHOUT a GPU (development tool, never imported by the
package, the tests or bench.py).

The C-ABI library is replaced by a stand-in whose entry points fill their outputs with cheap synthetic values of the
right shape (ratios around 1, z-scores around 0, one to three segments per chromosome): the NUMBERS ARE MEANINGLESS,
only the time the Python / NumPy side spends around the device calls is of interest (wall-clock of
a batch of 96 samples against the kernel time).  Time inside the stand-ins is reported separately and subtracted.

    python tools/host_profile_mock.py [batch] [--profile]
"""
import cProfile
import ctypes
import os
import pstats
import sys
import time
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wisecondorx_b200 import _lib, predict_control, predict_tools, synth  # noqa: E402


sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from fake_cabi import FakeLib as MockLib, make_ref_file  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 96
    mock = MockLib()
    _lib._lib = mock
    ref_file, _ = make_ref_file()
    samples, _ = synth.make_samples(batch, 15000, seed=3)
    args = types.SimpleNamespace(maskrepeats=5, minrefbins=150, alpha=1e-4, seed=1, gender=None, blacklist=None, zscore=5, beta=None)
    eng = predict_tools.PredictEngine(0)
    predict_control.predict_batch(args, samples[:1], [15000], ref_file, eng)
    for rep in range(2):
        mock.t = 0.0
        tim = {}
        prof = cProfile.Profile() if "--profile" in sys.argv and rep == 1 else None
        t0 = time.perf_counter()
        if prof:
            prof.enable()
        predict_control.predict_batch(args, samples, [15000] * batch, ref_file, eng, tim)
        if prof:
            prof.disable()
        wall = time.perf_counter() - t0
        print(f"batch {batch}: host wall {wall * 1e3:.0f} ms (stand-in library {mock.t * 1e3:.0f} ms inside), "
              f"normalize_and_assemble {tim['normalize_and_assemble'] * 1e3:.0f} ms, cbs_and_segment_z {tim['cbs_and_segment_z'] * 1e3:.0f} ms")
        if prof:
            pstats.Stats(prof).sort_stats("cumulative").print_stats(40)


if __name__ == "__main__":
    main()
