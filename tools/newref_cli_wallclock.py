#!/usr/bin/env python
"""Wall-clock of the drop-in `WisecondorX newref` command line (BASELINE.json metric, first half) on synthetic
sample files: writes S sample .npz files with the reference's schema, runs `python -m wisecondorx_b200.main newref`
in-process and prints the stage breakdown as one JSON line.  Optionally `predict` on one more sample."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wisecondorx_b200 import main as wmain, synth  # noqa: E402


def run():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=500)
    ap.add_argument("--binsize", type=int, default=15000)
    ap.add_argument("--predict", action="store_true")
    ap.add_argument("--cprofile", action="store_true", help="print the cumulative host profile of the newref call to stderr")
    ap.add_argument("--keep", default=None, help="directory to keep the files in (default: a temp dir)")
    a = ap.parse_args()
    d = a.keep or tempfile.mkdtemp(prefix="wcx_cli_")
    os.makedirs(d, exist_ok=True)
    t0 = time.perf_counter()
    cnv = [(a.samples, 5, 2000, 2400, 1.5)] if a.predict else None
    samples, genders = synth.make_samples(a.samples + (1 if a.predict else 0), a.binsize, seed=3, cnv=cnv)
    paths = []
    for i, s in enumerate(samples):
        p = os.path.join(d, "s%04d.npz" % i)
        np.savez_compressed(p, binsize=a.binsize, sample=s, quality={})
        paths.append(p)
    t_gen = time.perf_counter() - t0
    ref = os.path.join(d, "reference.npz")
    t0 = time.perf_counter()
    args = wmain.build_parser().parse_args(["newref"] + paths[: a.samples] + [ref, "--binsize", str(a.binsize), "--yfrac", "0.006"])
    if a.cprofile:
        import cProfile
        import pstats
        pr = cProfile.Profile()
        timings = pr.runcall(args.func, args)
        t_newref = time.perf_counter() - t0
        pstats.Stats(pr, stream=sys.stderr).sort_stats("cumulative").print_stats(45)
    else:
        timings = args.func(args)
        t_newref = time.perf_counter() - t0
    out = {"samples": a.samples, "binsize": a.binsize, "generate_inputs_s": round(t_gen, 2), "newref_wall_s": round(t_newref, 2),
           "stages_s": {k: round(v, 3) for k, v in (timings or {}).items()}, "reference_npz_mb": os.path.getsize(ref) >> 20}
    if a.predict:
        t0 = time.perf_counter()
        pargs = wmain.build_parser().parse_args(["predict", paths[-1], ref, os.path.join(d, "out"), "--bed"])
        res = pargs.func(pargs)
        out["predict_wall_s"] = round(time.perf_counter() - t0, 2)
        out["predict_stages_s"] = {k: round(v, 3) for k, v in res.get("timings", {}).items()}
        out["predict_segments"] = len(res["results_c"])
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    run()
