#!/usr/bin/env python
"""bench.py -- newref hot path (get_reference: all-pairs bin distance + top-refsize reference bins
+ null ratios) on synthetic input, BASELINE.json metric "bin-pair dist/s".

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload config3]

One "step" = one full pass of get_reference over every target bin of the workload.  For N > 1
(launched with torchrun, one rank per GPU) the target-bin axis is split with the reference's own
_get_part arithmetic (newref_tools.py:244-247); total work is fixed ("strong" scaling).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (config index for bin layout, samples, description)
    "config1": (1, 20, "newref 20 samples @ 1 Mb bins (2887 autosomal bins), refsize 300"),
    "config2": (2, 100, "newref 100 samples @ 100 kb bins (28760 autosomal bins), refsize 300"),
    "config3": (3, 500, "newref 500 samples @ 15 kb bins (191678 autosomal bins), refsize 300"),
}
REFSIZE = 300
NULL_M = 100


def n_pairs(per, row_begin=0, row_end=None):
    """bin-pair distances of the rows [row_begin,row_end): sum over rows of (N - n_chr(row))."""
    per = np.asarray(per, dtype=np.int64)
    cum = np.cumsum(per)
    n = int(cum[-1])
    row_end = n if row_end is None else row_end
    total = 0
    for c in range(len(per)):
        cs, ce = int(cum[c] - per[c]), int(cum[c])
        lo, hi = max(cs, row_begin), min(ce, row_end)
        if hi > lo:
            total += (hi - lo) * (n - int(per[c]))
    return total


def workload_config(name, n, s, pairs_total):
    """The `config` object: the workload only, identical in both arms (implementation notes go under `notes`)."""
    return {"workload": WORKLOADS[name][2], "refsize": REFSIZE, "null_samples": min(int(s), NULL_M), "bins": int(n),
            "samples": int(s), "pairs_per_step": int(pairs_total), "l2": "inputs larger than L2, no flush" if n * s * 8 > 200e6
            else "inputs fit in L2 (parity-size workload)"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons during the timed region."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []
        self.reasons = set()
        self.sm_max = None
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().splitlines()[0]
                f = [x.strip() for x in out.split(",")]
                self.samples.append(float(f[0]))
                self.sm_max = float(f[1])
                for nm, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def make_workload(name: str):
    from wisecondorx_b200 import synth
    cfg, s, _ = WORKLOADS[name]
    per = synth.config_bins(cfg)
    x, per, cum = synth.make_corrected_matrix(per, s, seed=cfg)
    return x, per, cum


def cpu_reference_sample(x, per, cum, target_seconds=20.0, check=None, max_windows=64):
    """Times the CPU restatement of the reference path (oracle/wcx_oracle.c, OpenMP over all host
    cores) on a bounded sample of target bins spread over the genome; returns pairs/s.
    check = (row_begin, idx, dist, null_ratios): host arrays of the GPU result for rows row_begin.. -- every timed
    window that lies inside them is compared with the oracle's output (indexes and distances bit-exact, null ratios
    1e-12) and the verdict is returned under "parity"."""
    from oracle import c_oracle
    c_oracle.build()
    n = x.shape[0]
    threads = max(c_oracle.max_threads(), len(os.sched_getaffinity(0)))  # torchrun exports OMP_NUM_THREADS=1: ask for the cores explicitly
    rng = np.random.default_rng(0)
    ids = list(range(min(x.shape[1], NULL_M)))
    # calibrate with one row per thread, then size the sample
    rows_done, pairs_done, t_total = 0, 0, 0.0
    batch = threads
    lo_s, hi_s = (0, n - batch) if check is None else (check[0], check[0] + check[1].shape[0] - batch)
    starts = lo_s + rng.permutation(max(1, hi_s - lo_s))[:max_windows]
    si = 0
    parity = {"rows": 0, "ok": True, "windows": 0, "checks": "indexes == , distances == (bit-exact), null ratios rtol 1e-12"}
    while t_total < target_seconds and si < len(starts):
        s0 = int(starts[si]); si += 1
        e0 = min(n, s0 + batch)
        t0 = time.perf_counter()
        idx, dist = c_oracle.topk(x, per, cum, REFSIZE, s0, e0, threads)
        nr = c_oracle.null_ratios(x, idx, s0, e0, ids, threads)
        t_total += time.perf_counter() - t0
        rows_done += e0 - s0
        pairs_done += n_pairs(per, s0, e0)
        if check is not None:
            rb, gi, gd, gn = check
            lo, hi = max(s0, rb), min(e0, rb + gi.shape[0])
            if hi > lo:
                a, b = lo - rb, hi - rb
                same = (np.array_equal(gi[a:b], idx[lo - s0:hi - s0]) and np.array_equal(gd[a:b], dist[lo - s0:hi - s0]) and
                        np.allclose(gn[a:b], nr[lo - s0:hi - s0], rtol=1e-12, atol=1e-14, equal_nan=True))
                parity["rows"] += hi - lo
                parity["windows"] += 1
                if not same:
                    parity["ok"] = False
                    parity.setdefault("first_mismatch_row", lo)
        if si == 1 and t_total > 0:
            est = target_seconds / t_total
            batch = int(max(threads, min(4096, batch * max(1.0, est / 8))))
    out = {"value": pairs_done / t_total, "unit": "bin-pair dist/s", "cores": threads, "kind": "port",
           "sample": f"{rows_done} target bins in {si} windows spread over the genome (all {n} candidates each), "
                     f"C restatement of get_reference with OpenMP, {t_total:.1f} s"}
    if check is not None:
        out["parity"] = parity
    return out


def cli_and_predict_extras(device, n_train=500, binsize=15000, batch=96, cpu_baseline=True, batches=None):
    """The rest of the BASELINE metric, measured in the same run on the same box:

    * `cli`: wall-clock of the drop-in command line at config 3 -- `WisecondorX newref` (A + F + M passes, .npz in,
      reference .npz out) on 500 synthetic sample files at 15 kb, then `WisecondorX predict --bed` of one more sample
      against the reference it wrote.  The sample files are generated and written outside the timed regions.
    * `predict` (configs 4 / 5): the whole numeric flow of predict (predict_control.predict_batch: gender, both
      `normalize` calls, assembly, log transform, CBS through exec_cbs, segment z-scores) for 1 and 96 samples against
      that reference, with the device time of every kernel family, a roofline entry for the normalisation passes and
      the CPU baseline: the NumPy restatement of `normalize` x2 + `get_z_score` (oracle/np_oracle.py, kind "port",
      pinned to the live reference by tests/test_oracle_pin.py) on one sample, one core, at the same size."""
    import shutil
    import tempfile
    import types
    from concurrent.futures import ThreadPoolExecutor
    from wisecondorx_b200 import _lib, main as wmain, npz_io, predict_control, predict_tools, synth
    batches = batches or (1, batch)
    d = tempfile.mkdtemp(prefix="wcx_bench_")
    out = {}
    try:
        t0 = time.perf_counter()
        samples, genders = synth.make_samples(n_train + batch, binsize, seed=3, cnv=[(n_train, 5, 2000, 2400, 1.5)])
        paths = [os.path.join(d, "s%04d.npz" % i) for i in range(n_train + 1)]
        with ThreadPoolExecutor(min(32, len(os.sched_getaffinity(0)))) as pool:
            list(pool.map(lambda i: np.savez_compressed(paths[i], binsize=binsize, sample=samples[i], quality={}), range(n_train + 1)))
        t_gen = time.perf_counter() - t0
        ref = os.path.join(d, "reference.npz")
        parser = wmain.build_parser()
        t0 = time.perf_counter()
        a = parser.parse_args(["newref"] + paths[:n_train] + [ref, "--binsize", str(binsize), "--yfrac", "0.006", "--device", str(device)])
        st_new = a.func(a)
        t_newref = time.perf_counter() - t0
        t0 = time.perf_counter()
        a = parser.parse_args(["predict", paths[n_train], ref, os.path.join(d, "out"), "--bed", "--seed", "1", "--device", str(device)])
        res = a.func(a)
        t_predict = time.perf_counter() - t0
        ab = [l.split("\t") for l in open(os.path.join(d, "out_aberrations.bed")).read().splitlines()[1:]]
        out["cli"] = {"config": "500 samples @ 15 kb (A + F + M passes), refsize 300; predict of 1 sample with a planted chr5 gain",
                      "generate_and_write_inputs_s": round(t_gen, 2), "newref_wall_s": round(t_newref, 2),
                      "newref_stages_s": {k: round(v, 3) for k, v in (st_new or {}).items()},
                      "reference_npz_mb": os.path.getsize(ref) >> 20, "predict_wall_s": round(t_predict, 2),
                      "predict_stages_s": {k: round(v, 3) for k, v in res.get("timings", {}).items()},
                      "predict_segments": len(res["results_c"]),
                      "planted_gain_found": any(x[0] == "5" and x[5] == "gain" for x in ab)}
        # ---- predict batches through the library flow
        out["predict"] = {"error": "not finished"}  # replaced below; what the line shows if this part raises
        ref_file = npz_io.load_npz(ref)
        eng = predict_tools.PredictEngine(device)
        args = types.SimpleNamespace(maskrepeats=5, minrefbins=150, alpha=1e-4, seed=1, gender=None, blacklist=None, zscore=5, beta=None)
        tests = samples[n_train:]
        predict_control.predict_batch(args, tests[:1], [binsize], ref_file, eng)  # reference arrays to the device, warm-up
        n_a = int(ref_file["masked_bins_per_chr_cum"][-1])
        k = int(ref_file["indexes"].shape[1])
        pred = {}
        for b in batches:
            first_ms = None
            if b > 1:
                # first call of a batch size: page-locked staging buffers are allocated (cudaHostAlloc, ~0.2 ms / MB);
                # they are recycled by every later call (wisecondorx_b200/_lib.py PinnedPool): the steady state is timed
                t0 = time.perf_counter()
                predict_batch_timed(args, tests, b, binsize, ref_file, eng, {})
                first_ms = (time.perf_counter() - t0) * 1e3
            eng.ctx.__dict__["kernel_ms_acc"] = {}
            tim = {}
            t0 = time.perf_counter()
            outs = predict_batch_timed(args, tests, b, binsize, ref_file, eng, tim)
            wall = time.perf_counter() - t0
            acc = dict(eng.ctx.__dict__["kernel_ms_acc"])
            nk = acc.get("coverage_project", 0) + acc.get("gather_list", 0) + acc.get("passes", 0) + acc.get("medians", 0)
            pred[f"batch{b}"] = {"wall_ms": wall * 1e3, "first_call_wall_ms": first_ms, "normalize_and_assemble_wall_ms": tim["normalize_and_assemble"] * 1e3,
                                 "cbs_and_segment_z_wall_ms": tim["cbs_and_segment_z"] * 1e3, "normalize_kernels_ms": nk,
                                 "cbs_kernels_ms": acc.get("cbs", 0.0), "segment_z_kernels_ms": acc.get("segment_z", 0.0),
                                 "kernels_ms": {kk: round(v, 3) for kk, v in acc.items()},
                                 "segments": int(sum(len(o[1]["results_c"]) for o in outs)), "cbs_stats": cbs_stats(eng)}
        # roofline of the dominant predict kernel family: the three passes of the AUTOSOMAL normalize of one sample,
        # algorithmic bytes per SURVEY.md 8(d): 3 * N * k * (4 + 8) + 3 * 4 * N * 8
        eng.ctx.__dict__["kernel_ms_acc"] = {}
        predict_control.normalize_batch(args, [predict_control.resolve_genders(args, dict(tests[0]), ref_file)[0]], ref_file, "A", eng)
        pm = eng.ctx.__dict__["kernel_ms_acc"]
        bytes_k9 = 3.0 * n_a * k * 12 + 3.0 * 4 * n_a * 8
        peaks = load_peaks()
        ach = bytes_k9 / (pm["passes"] * 1e-3) / 1e9
        pred["roofline"] = {"bound": "hbm", "kernel": "normalize_pass_kernel x3 (autosomal normalize, batch 1)", "achieved": ach,
                            "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "kernel_ms": pm["passes"],
                            "algorithmic_bytes": bytes_k9, "medians_ms": pm["medians"], "coverage_project_ms": pm["coverage_project"],
                            "note": "bytes_K9 of SURVEY.md 8(d); the kernels read a third of it (4-byte gather lists resolved once per "
                                    "(reference, cutoff) instead of idx + dist rows in every pass)",
                            "peak_src": peaks["src"] + ": hbm_gbs"}
        if cpu_baseline:
            pred["cpu_baseline"] = predict_cpu_baseline(tests[0], ref_file, outs[0] if batches[-1] == 1 else None, args)
        out["predict"] = pred
    except Exception as e:  # the parts measured so far stay in the line
        import traceback
        traceback.print_exc()
        out.setdefault("cli", {"error": repr(e)})
        if "error" in out.get("predict", {"error": 1}):
            out["predict"] = {"error": repr(e)}
    finally:
        shutil.rmtree(d, ignore_errors=True)
    return out


def predict_batch_timed(args, tests, b, binsize, ref_file, eng, tim):
    from wisecondorx_b200 import predict_control
    return predict_control.predict_batch(args, tests[:b], [binsize] * b, ref_file, eng, tim)


def cbs_stats(eng):
    from wisecondorx_b200 import cbs
    return cbs.cbs_stats(eng.ctx)


def predict_cpu_baseline(sample, ref_file, gpu_out, args):
    """The reference's predict numeric path on the host: normalize (predict_control.py:21-39) for the autosomes and the
    gonosomes + get_z_score (overall_tools.py:88-119) over whole-chromosome segments, restated in NumPy
    (oracle/np_oracle.py), one sample, one core, at 15 kb against the reference the GPU just built.  CBS is excluded on
    the CPU side (R / DNAcopy are not installed)."""
    from oracle import np_oracle
    from wisecondorx_b200 import predict_control
    s2, gender, rg = predict_control.resolve_genders(args, dict(sample), ref_file)
    t0 = time.perf_counter()
    aut = np_oracle.normalize(s2, ref_file, "A", args.maskrepeats)
    t_a = time.perf_counter() - t0
    t0 = time.perf_counter()
    gon = np_oracle.normalize(s2, ref_file, rg, args.maskrepeats)
    t_g = time.perf_counter() - t0
    return {"value": t_a + t_g, "unit": "s per sample (normalize x2)", "cores": 1, "kind": "port",
            "normalize_autosomes_s": t_a, "normalize_gonosomes_s": t_g,
            "sample": "np_oracle.normalize for the autosomal and the gonosomal reference of one 15 kb sample (all bins), "
                      "NumPy single process; CBS excluded (no R)"}


def predict_sharded_extras(eng, x, per, cum, idx_full, dist_full, device, rank, world):
    """BASELINE config 5 on N GPUs: 96 test samples sharded over the ranks (reference arrays replicated, no
    collective on the data path); wall-clock = max over ranks of normalize + CBS for the rank's samples."""
    import types
    import torch
    import torch.distributed as dist
    from wisecondorx_b200 import parallel, predict_tools
    n, s = x.shape
    rng = np.random.default_rng(99)
    comps = np.linalg.qr(rng.standard_normal((n, 5)))[0].T.copy()
    ref = {"indexes": idx_full, "distances": dist_full, "masked_bins_per_chr": per, "masked_bins_per_chr_cum": cum,
           "pca_components": comps, "pca_mean": np.full(n, 1.0 / n), "mask": np.ones(n, dtype=bool), "bins_per_chr": per}
    offs = np.concatenate([[0], cum]).astype(int)
    a, b = parallel.shard_samples(96, world)[rank]
    samples = [None] * 96
    for i in range(96):  # same stream on every rank; only the rank's own samples are kept
        lam = 60.0 * np.clip(x[:, i % s], 0, None)
        lam[offs[4] + 2000: offs[4] + 2400] *= 1.5
        c = rng.poisson(lam).astype(np.int32)
        if a <= i < b:
            samples[i] = {str(k + 1): c[offs[k]:offs[k + 1]] for k in range(22)}
    args = types.SimpleNamespace(maskrepeats=5, minrefbins=150, alpha=1e-4, seed=1)
    pe = predict_tools.PredictEngine(device, eng.ctx)
    parallel.predict_batch_sharded(args, samples, ref, "A", pe)  # warm-up: reference arrays to the device, staging buffers allocated
    dist.barrier()
    t0 = time.perf_counter()
    (_, _), _, summary = parallel.predict_batch_sharded(args, samples, ref, "A", pe)
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=torch.device("cuda", device))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"samples": 96, "samples_per_rank": b - a, "wall_ms_max_over_ranks": float(t.item()) * 1e3,
            "segments": int(sum(summary)) if summary is not None else None}


def _cbs_ms(eng):
    import ctypes
    from wisecondorx_b200 import _lib
    o = np.zeros(4)
    _lib.check(_lib.load().wcx_predict_stage_ms(eng.ctx.handle, ctypes.c_void_p(o.ctypes.data)))
    return float(o[3])


def run_reference(args, rank, world):
    if rank != 0:
        return
    x, per, cum = make_workload(args.workload)
    per_step = max(3.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_reference_sample(x, per, cum, per_step)
        if i >= args.warmup:
            vals.append(r)
    v = float(np.mean([r["value"] for r in vals]))
    total_pairs = n_pairs(per)
    out = {
        "impl": "reference", "metric": "newref bin-pair distances per second", "value": v, "unit": "bin-pair dist/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_pairs / v * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, x.shape[0], x.shape[1], total_pairs),
        "notes": {"ms_per_step": "extrapolated from the bounded sample"},
        "cpu_baseline": {"value": v, "unit": "bin-pair dist/s", "cores": vals[-1]["cores"], "kind": "port",
                         "sample": vals[-1]["sample"]},
        "e2e": {"value": v, "unit": "bin-pair dist/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from wisecondorx_b200 import _lib, newref_tools

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    x, per, cum = make_workload(args.workload)
    n, s = x.shape
    rb, re = newref_tools._get_part(rank, world, n)
    rows = re - rb
    k = REFSIZE
    m = min(s, NULL_M)
    ids = np.arange(m, dtype=np.int32)  # fixed null-sample columns (the reference draws them at random)
    pairs_total = n_pairs(per)

    eng = newref_tools.NewrefEngine(local_rank, _lib.Context(local_rank))
    stream = torch.cuda.current_stream(dev)
    eng.ctx.set_stream(stream.cuda_stream)
    # ---- device-resident arm ("value"): X already in HBM, outputs stay in HBM
    x_dev = torch.from_numpy(x).to(dev)
    eng.load(None, per, cum, on_device_ptr=x_dev.data_ptr(), shape=(n, s))
    idx_dev = torch.empty((rows, k), dtype=torch.int32, device=dev)
    dist_dev = torch.empty((rows, k), dtype=torch.float64, device=dev)
    nr_dev = torch.empty((rows, m), dtype=torch.float64, device=dev)

    def step_resident():
        if args.unfused:
            eng.topk(rb, re, k, kernel=args.kernel, device_out=(idx_dev.data_ptr(), dist_dev.data_ptr()))
            eng.null_ratios(rb, re, k, ids, device_out=nr_dev.data_ptr())
        else:  # one C-ABI call: sweep, then re-rank blocks with the null-ratio kernels of finished blocks on a side stream
            eng.reference(rb, re, k, ids, kernel=args.kernel, device_out=(idx_dev.data_ptr(), dist_dev.data_ptr(), nr_dev.data_ptr()))

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup, sample_clocks=False):
        for _ in range(warmup):
            fn()
        barrier()
        sampler = None
        if sample_clocks and rank == 0:
            sampler = ClockSampler(local_rank)
            sampler.start()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        extra = []
        for _ in range(steps):
            extra.append(fn())
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), clocks, extra

    launches0 = eng.stats()["launches"]

    stage_acc = {"sweep": 0.0, "rerank": 0.0, "exact_rows": 0.0, "null_ratios": 0.0, "sweep_tail_past_main": 0.0}

    def step_resident_timed():
        step_resident()
        sm = eng.stage_ms()
        for kk in stage_acc:
            stage_acc[kk] += sm[kk]

    ms_total, clocks, _ = timed(step_resident_timed, args.steps, args.warmup, sample_clocks=True)
    for kk in stage_acc:  # warm-up steps were accumulated too: rescale to the timed ones
        stage_acc[kk] *= args.steps / float(args.steps + args.warmup)
    st = eng.stats()
    launches = (st["launches"] - launches0) * args.steps // (args.steps + args.warmup)
    ms_per_step = ms_total / args.steps
    value = pairs_total / (ms_per_step * 1e-3)

    # ---- end-to-end arm: host (pinned) buffers through the C-ABI, H2D + D2H inside the timed region
    if world == 1:
        x_pin = torch.from_numpy(x).pin_memory()
        idx_pin = torch.empty((rows, k), dtype=torch.int32).pin_memory()
        dist_pin = torch.empty((rows, k), dtype=torch.float64).pin_memory()
        nr_pin = torch.empty((rows, m), dtype=torch.float64).pin_memory()
        xh, ih, dh, nh = x_pin.numpy(), idx_pin.numpy(), dist_pin.numpy(), nr_pin.numpy()

        def step_e2e():
            # one C-ABI call with host buffers: H2D of X, all kernels, D2H of the three outputs
            eng.get_reference_host(xh, per, cum, rb, re, k, ids, kernel=args.kernel, out=(ih, dh, nh))

        h2d = x.nbytes
        d2h = ih.nbytes + dh.nbytes + nh.nbytes
        e2e_ms, _, _ = timed(step_e2e, max(1, args.steps // 2), max(1, args.warmup))
        e2e_ms /= max(1, args.steps // 2)
        # self-check: the host-to-host call and the device-resident call return the same arrays
        e2e_same = bool(np.array_equal(ih, idx_dev.cpu().numpy()) and np.array_equal(dh, dist_dev.cpu().numpy())
                        and np.array_equal(nh, nr_dev.cpu().numpy(), equal_nan=True))
        e2e_path = "wcx_get_reference: pinned host X in, pinned host (indexes, distances, null ratios) out"
    else:
        # every rank uploads its slice of X, one NCCL all-gather over NVLink gives every GPU the matrix, sharded
        # compute, every rank writes its row block into one shared page-locked host segment (parallel.ShardedReference)
        from wisecondorx_b200 import parallel
        h2d = x.nbytes
        d2h = n * k * 12 + n * m * 8
        try:
            sr = parallel.ShardedReference(n, s, k, m, dev, engine=eng)
        except RuntimeError:
            sr = None  # no room for the shared segment (raised on every rank): gather through rank 0 instead
        if sr is not None:
            x_slice_pin = torch.from_numpy(np.ascontiguousarray(sr.slice_of(x))).pin_memory()

            def step_e2e():
                sr.run(x_slice_pin, per, cum, ids)

            e2e_ms, _, _ = timed(step_e2e, max(1, args.steps // 2), max(1, args.warmup))
            e2e_phases = {}
            sr.run(x_slice_pin, per, cum, ids, phases=e2e_phases)  # one more pass with a synchronisation after every phase
            ho = sr.host_out
            e2e_same = bool(np.array_equal(ho[0][rb:re], idx_dev.cpu().numpy()) and np.array_equal(ho[1][rb:re], dist_dev.cpu().numpy())
                            and np.array_equal(ho[2][rb:re], nr_dev.cpu().numpy(), equal_nan=True))
            predict_sharded = None
            if not args.no_predict:
                try:  # an extra: its failure must not cost the headline line (every rank runs the same host code)
                    predict_sharded = predict_sharded_extras(eng, x, per, cum, np.array(ho[0]), np.array(ho[1]), local_rank, rank, world)
                except Exception as e:
                    import traceback
                    traceback.print_exc()
                    predict_sharded = {"error": repr(e)}
            del ho
            sr.close()
            e2e_path = "sliced H2D + NCCL all-gather of X, row blocks written by every rank into one shared pinned host segment"
        else:
            holder = {}

            def step_e2e():
                holder["out"] = parallel.get_reference_sharded(x if rank == 0 else None, per, cum, k, ids, device=dev, engine=eng)

            e2e_ms, _, _ = timed(step_e2e, max(1, args.steps // 2), max(1, args.warmup))
            e2e_same = True
            if rank == 0:
                o = holder["out"]
                e2e_same = bool(np.array_equal(o[0][rb:re], idx_dev.cpu().numpy()) and np.array_equal(o[1][rb:re], dist_dev.cpu().numpy()))
            e2e_path = "H2D on rank 0 + NCCL broadcast of X, gather of the row blocks to rank 0, D2H on rank 0"
            predict_sharded = None
        e2e_ms /= max(1, args.steps // 2)
        flag = torch.tensor([1 if e2e_same else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # every rank checks its own block of the host arrays
        e2e_same = bool(flag.item())
    e2e_value = pairs_total / (e2e_ms * 1e-3)
    e2e_phases = locals().get("e2e_phases")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    idx_host, dist_host, nr_host = idx_dev.cpu().numpy(), dist_dev.cpu().numpy(), nr_dev.cpu().numpy()
    peaks = load_peaks()
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["dist_topk_tc_kernel"]
        if tj["workload"] == args.workload and tj["n_gpus"] == world:
            traffic = tj["dram_bytes_per_launch"]
    except Exception:
        pass
    # dominant kernel: the tensor-core sweep.  Algorithmic flops = 2 * S * pairs;
    # TF32 dense peak taken as half the measured bf16 cuBLAS figure (nominal 1:2 ratio).
    # The sweep is two launches of the same kernel: the main one (whole rounds of the persistent grid, all SMs) and the
    # partial last round (its own launch beside the re-rank).  The roofline entry is the MAIN launch: its flops (the rows
    # it sweeps) over its own duration (stages_ms.sweep minus what the partial round runs past it).
    rows_main = st.get("rows_main_sweep", rows) or rows
    my_pairs = n_pairs(per, rb, rb + rows_main)
    sweep_ms = (stage_acc["sweep"] - stage_acc["sweep_tail_past_main"]) / args.steps
    achieved = 2.0 * s * my_pairs / (sweep_ms * 1e-3) / 1e12 if sweep_ms > 0 else None
    # operand type of the sweep actually run: f16 (default, kernels 0 / 5 / 6) runs at the bf16/f16 tensor rate the
    # driver measured; tf32 (kernels 1 / 4) at half of it (nominal ratio)
    f16 = args.kernel in (0, 5, 6)
    peak_tc = peaks["bf16_tflops_sustained"] if f16 else peaks["bf16_tflops_sustained"] / 2.0
    out = {
        "metric": "newref bin-pair distances per second", "value": value, "unit": "bin-pair dist/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args.workload, n, s, pairs_total),
        "notes": {"arithmetic": "results are float64, bit-exact with the reference's NumPy order (exact re-rank); the sweep that "
                                "nominates candidates runs on " + ("f16" if f16 else "tf32") + " tensor-core operands with fp32 accumulation",
                  "l2": ("inputs (X fp64 %.0f MB + operands) larger than the 126 MB L2, no flush needed" % (x.nbytes / 1e6)) if x.nbytes > 200e6
                        else "inputs fit in L2 and are not flushed: parity-size workload, not the bench line",
                  "parallelism": f"target-bin parts x{world}",
                  "null_ratios": "separate call after the top-k" if args.unfused else
                                 "resident outputs: one launch after the re-rank of each region (inside stages_ms.rerank); "
                                 "host outputs (e2e): row blocks on a side stream next to the re-rank of the following block"},
        "e2e": {"value": e2e_value, "unit": "bin-pair dist/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms, "equals_resident_result": e2e_same, "path": e2e_path,
                "phases_ms_cumulative_rank0": {k: round(v, 2) for k, v in e2e_phases.items()} if e2e_phases else None},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "dist_topk_tc_kernel", "achieved": achieved, "peak": peak_tc,
                     "unit": "TFLOP/s", "frac": (achieved / peak_tc) if achieved else None, "traffic": traffic,
                     "peak_src": f"{peaks['src']}: bf16_tflops_sustained" + ("" if f16 else " / 2 (tf32)"),
                     "kernel_ms": sweep_ms, "rows_of_launch": int(rows_main), "rows_total": int(rows)},
        "stages_ms": {kk: v / args.steps for kk, v in stage_acc.items()},
        "exact_fallback_rows": st["exact_fallback_rows"],
        "counters": {"rerank_exact_distances_per_row": eng.stage_ms()["exact_evals"] / max(1, rows),
                     "listed_candidates_per_row": eng.stage_ms()["gathered_entries"] / max(1, rows)},
    }
    if world == 1 and not args.no_predict:
        try:  # extras beside the headline: a failure is reported in their place, not instead of the line
            out.update(cli_and_predict_extras(local_rank, cpu_baseline=not args.no_cpu_baseline))
        except (Exception, SystemExit) as e:
            import traceback
            traceback.print_exc()
            out.setdefault("cli", {"error": repr(e)})
            out.setdefault("predict", {"error": repr(e)})
    if world > 1 and predict_sharded is not None:
        out["predict"] = {"batch96_sharded": predict_sharded}
    # parity of THIS run's result at THIS configuration: the oracle's rows against the GPU arrays (the CPU baseline's
    # timed windows double as the parity sample; without the baseline leg a short sample is still checked)
    check = (rb, idx_host, dist_host, nr_host)
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_reference_sample(x, per, cum, args.cpu_seconds, check=check)
    else:
        cb = cpu_reference_sample(x[:, :], per, cum, 2.0, check=check, max_windows=3)
    out["parity"] = cb.pop("parity")
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cb
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if not out["parity"]["ok"] or out["parity"]["rows"] == 0:
        sys.stderr.write("bench.py: GPU result differs from the oracle (or nothing was compared): %r\n" % (out["parity"],))
        sys.exit(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-predict", action="store_true", help="skip the predict (configs 4 / 5) timings")
    ap.add_argument("--kernel", type=int, default=0, help="sweep kernel (include/wcx_b200.h WCX_KERNEL_*), 0 = default")
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    ap.add_argument("--unfused", action="store_true", help="separate topk and null-ratio calls (stage timings of the null kernel)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
