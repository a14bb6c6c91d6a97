"""Host-side profile of `predict_control.predict_batch` WITHOUT a GPU (development tool, never imported by the
package, the tests or bench.py).

The C-ABI library is replaced by a stand-in whose entry points fill their outputs with cheap synthetic values of the
right shape (ratios around 1, z-scores around 0, one to three segments per chromosome): the NUMBERS ARE MEANINGLESS,
only the time the Python / NumPy side spends around the device calls is of interest (wall-clock of
a batch of 96 samples against the kernel time).  Time inside the stand-ins is reported separately and subtracted.

    python tools/host_profile_mock.py [batch] [--profile]      host side of predict_batch
    python tools/host_profile_mock.py newref [samples]         `WisecondorX newref` at 15 kb, stages as the command line logs them
                                                               (get_reference.* is the stand-in filling fresh pages, not device time)
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


def newref(n_samples):
    import shutil
    import tempfile
    from concurrent.futures import ThreadPoolExecutor
    from wisecondorx_b200 import main as wmain
    _lib._lib = MockLib()
    d = tempfile.mkdtemp(prefix="wcx_host_profile_")
    try:
        samples, _ = synth.make_samples(n_samples, 15000, seed=3)
        paths = [os.path.join(d, "s%04d.npz" % i) for i in range(n_samples)]
        with ThreadPoolExecutor(8) as pool:
            list(pool.map(lambda i: np.savez_compressed(paths[i], binsize=15000, sample=samples[i], quality={}), range(n_samples)))
        del samples
        parser = wmain.build_parser()
        for rep in range(2):
            a = parser.parse_args(["newref"] + paths + [os.path.join(d, "ref.npz"), "--binsize", "15000", "--yfrac", "0.006"])
            t0 = time.perf_counter()
            stages = a.func(a)
            print(f"newref, {n_samples} samples: host wall {time.perf_counter() - t0:.2f} s", {k: round(v, 3) for k, v in stages.items()})
    finally:
        shutil.rmtree(d, ignore_errors=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "newref":
        return newref(int(sys.argv[2]) if len(sys.argv) > 2 else 500)
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
