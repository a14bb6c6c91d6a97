#!/usr/bin/env python
"""The predict workload the ncu captures / host profiles are taken on: `WisecondorX newref` on 500 synthetic samples at
15 kb (so the reference is the real thing: A / F / M sets, PCA model, masks, null ratios), then the library's predict flow
(predict_control.predict_batch: both normalize calls, assembly, CBS, segment z-scores) for a batch of samples.

    python tools/predict_profile.py [--batch 1] [--cprofile]      (tools/gpu_run.sh ncu_predict / prof_batch)
"""
import argparse
import cProfile
import io
import json
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--train", type=int, default=500)
    ap.add_argument("--cprofile", action="store_true")
    a = ap.parse_args()
    if a.cprofile:
        pr = cProfile.Profile()
        orig = bench.predict_batch_timed

        def wrapped(*args, **kw):
            if args[2] >= a.batch and a.batch > 1:
                pr.enable()
                try:
                    return orig(*args, **kw)
                finally:
                    pr.disable()
            return orig(*args, **kw)

        bench.predict_batch_timed = wrapped
    out = bench.cli_and_predict_extras(0, n_train=a.train, batch=max(a.batch, 1), cpu_baseline=False, batches=(1, a.batch) if a.batch > 1 else (1,))
    print(json.dumps(out["predict"]))
    if a.cprofile:
        sio = io.StringIO()
        pstats.Stats(pr, stream=sio).sort_stats("cumulative").print_stats(45)
        print(sio.getvalue())


if __name__ == "__main__":
    main()
