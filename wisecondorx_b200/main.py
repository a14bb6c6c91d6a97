"""`WisecondorX` command line on the B200 kernels: the sub-commands, flags, defaults and .npz / .bed
formats of the reference CLI (reference main.py:302-502) so that `newref` and `predict` are
drop-in.  `convert` (BAM/CRAM ingest, pysam) and `--plot` (R) are outside the accelerated path and
not provided here; use the reference for those steps -- the .npz files are interchangeable.

Error convention of the reference is kept: user errors log a critical message and `sys.exit()`
(status 0)."""
from __future__ import annotations

import argparse
import logging
import os
import sys
import warnings

import numpy as np

import time

from . import _lib, cbs, newref_control, newref_tools, npz_io, predict_control, predict_output, predict_tools, ref_qc
from .overall_tools import gender_correct, scale_sample


# ---------------------------------------------------------------------------------------------
# host-side helpers of newref (reference newref_tools.py:21-102) -- O(samples) / one pass, host
# ---------------------------------------------------------------------------------------------
def _y_fraction(sample):
    return float(np.sum(sample["24"])) / float(np.sum([np.sum(sample[x]) for x in sample.keys()]))


def train_gender_model(args, samples, y_fractions=None):
    """Two-component Gaussian mixture on the Y-read fraction; the first local minimum of the mixture
    density on [0, 0.02] is the male/female cutoff unless --yfrac is given (reference :21-68).
    y_fractions: the fractions when the caller has them (stacked_counts below)."""
    y = np.array([_y_fraction(s) for s in samples]) if y_fractions is None else np.asarray(y_fractions, dtype=float)
    if args.yfrac is not None:
        cut_off = args.yfrac
    else:
        from scipy.signal import argrelextrema
        from sklearn.mixture import GaussianMixture
        gmm = GaussianMixture(n_components=2, covariance_type="full", reg_covar=1e-99, max_iter=10000, tol=1e-99)
        gmm.fit(X=y.reshape(-1, 1))
        gx = np.linspace(0, 0.02, 5000)
        gy = np.exp(gmm.score_samples(gx.reshape(-1, 1)))
        cut_off = gx[argrelextrema(gy, np.less)][0]
        logging.info("Determined --yfrac cutoff: {}".format(str(round(cut_off, 4))))
    genders = np.empty(len(samples), dtype="object")
    genders[y > cut_off] = "M"
    genders[y < cut_off] = "F"
    return genders.tolist(), cut_off


class stacked_counts:
    """The read counts of all samples as one int32 matrix [bins, samples] (newref_tools.stack_counts) with the exact
    read totals per sample, built ONCE for the Y fractions of the gender model, the three coverage masks and the three
    passes of `newref`.  The reference recomputes all of these from the per-sample dicts (newref_tools.py:21-30,
    :77-102, :110-129); the integer sums are the same, so are the float64 quotients."""

    def __init__(self, samples):
        self.bins_per_chr = [max(len(s[str(c)]) for s in samples) for c in range(1, 25)]
        self.offs = np.concatenate([[0], np.cumsum(self.bins_per_chr)]).astype(np.int64)
        self.counts = newref_tools.stack_counts(list(samples), range(1, 25))
        self.totals = newref_tools.column_totals(self.counts)
        self.x_totals = newref_tools.column_totals(self.counts[self.offs[22]:self.offs[23]])
        self.y_totals = newref_tools.column_totals(self.counts[self.offs[23]:self.offs[24]])
        # the reference's denominator adds up EVERY key of the sample dict (predict_tools.py:17-24, newref_tools.py:24-27):
        # identical to the column total when the keys are the 24 chromosomes and no sample is empty
        self.y_fractions = None
        if all(len(s) == 24 for s in samples) and not np.any(self.totals == 0):
            self.y_fractions = self.y_totals.astype(float) / self.totals.astype(float)

    def gender_correct(self, genders):
        """overall_tools.gender_correct for every male column: X and Y counts doubled (reference overall_tools.py:48-53)."""
        male = np.flatnonzero(np.array(genders, dtype=object) == "M")
        if len(male):
            self.counts[self.offs[22]:self.offs[24], male] *= 2
            self.totals[male] += self.x_totals[male] + self.y_totals[male]
            self.x_totals[male] *= 2
            self.y_totals[male] *= 2


def get_mask(samples, counts=None, cols=None, totals=None):
    """Bins with more than 5 % of the median (non-zero) summed normalised coverage (reference newref_tools.py:77-102).
    Same arithmetic as the reference, element for element -- exact column totals (integers), one division per element,
    the additions of a bin in NumPy's pairwise order -- but the [bins, samples] float matrix (0.8 GB at 15 kb / 500
    samples) is never materialised: newref_tools.bin_sums walks the int32 count matrix on host threads.
    counts / cols: the stacked count matrix of a superset of `samples` with the same bins per chromosome and the columns
    of `samples` in it (tool_newref stacks the samples once for the three masks and the three passes); totals: the
    column totals of that matrix (newref_tools.column_totals) when the caller has them."""
    bins_per_chr = [max(len(s[str(c)]) for s in samples) for c in range(1, 25)]
    total = int(sum(bins_per_chr))
    if counts is None or counts.shape[0] != total:
        counts, cols, totals = newref_tools.stack_counts(samples, range(1, 25)), None, None  # int32 [total, S], zero padded
    if totals is None:
        totals = newref_tools.column_totals(counts)
    col_sum = np.asarray(totals, dtype=np.int64).astype(float)   # exact, like the float sum of integer counts
    if cols is not None:
        col_sum = col_sum[cols]
    sum_per_bin = newref_tools.bin_sums(counts, col_sum, cols)
    median_cov = np.median(sum_per_bin[sum_per_bin > 0])
    return sum_per_bin > (0.05 * median_cov), bins_per_chr


predict_gender = predict_control.predict_gender


# ---------------------------------------------------------------------------------------------
# newref
# ---------------------------------------------------------------------------------------------
def _prewarm_newref(args):
    """Starts, on a background thread, what the GPU passes would otherwise pay for on the critical path: the CUDA
    context and the page-locked result staging the A / F / M passes share (0.1-0.4 ms per MB, 0.9 GB at 15 kb).  Reading the
    samples and building the gender model and the masks keep the host busy for longer than that.  The sizes are upper
    bounds taken from the first sample file (unmasked bins after re-binning); a pass whose arrays turn out larger just
    allocates its own."""
    device = getattr(args, "device", 0)
    sizes = []
    try:
        with np.load(args.infiles[0], encoding="latin1", allow_pickle=True) as f:
            sample, from_size = f["sample"].item(), int(f["binsize"].item())
        scale = max(1, int(args.binsize // from_size)) if args.binsize else 1
        per = [-(-len(sample.get(str(c), ())) // scale) for c in range(1, 25)]
        m_null = min(len(args.infiles), 100)
        t = int(sum(per))  # ONE set, sized for the largest pass: every pass copies its results out of it (tool_newref)
        sizes = [t * args.refsize * 4, t * args.refsize * 8, t * m_null * 8]
    except Exception:  # unreadable first file: the regular loading code reports it
        sizes = []
    return _lib.prewarm_async(device, sizes)


def tool_newref(args):
    logging.info("Creating new reference")
    if args.yfrac is not None and (args.yfrac < 0 or args.yfrac > 1):
        logging.critical("Parameter --yfrac should be a positive number lower than or equal to 1")
        sys.exit()
    if getattr(args, "plotyfrac", None) is not None:
        # the reference draws the Y-fraction histogram with matplotlib and exits (newref_tools.py:44-58)
        logging.critical("--plotyfrac needs the reference's matplotlib plot and is not part of the accelerated path; "
                         "run the reference's `WisecondorX newref --plotyfrac` for the histogram")
        sys.exit()
    samples = []
    timings = {}
    t0 = time.perf_counter()
    _prewarm_newref(args)  # CUDA context + page-locked result buffers, beside the host-only stages that follow
    logging.info("Importing data ...")
    for infile, (sample, binsize) in zip(args.infiles, npz_io.load_samples(args.infiles)):  # inflated concurrently
        logging.info("Loading: {}".format(infile))
        samples.append(scale_sample(sample, binsize, args.binsize))
    samples = np.array(samples)
    timings["load_samples"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    # one stacked count matrix [bins, samples] for the Y fractions, the three masks and the three passes
    stacked = stacked_counts(samples) if len(samples) else None
    genders, trained_cutoff = train_gender_model(args, samples, stacked.y_fractions if stacked else None)
    if genders.count("F") < 5 and args.nipt:
        logging.warning("A NIPT reference should have at least 5 female feti samples. Removing --nipt flag.")
        args.nipt = False
    if not args.nipt:
        for i, sample in enumerate(samples):
            samples[i] = gender_correct(sample, genders[i])
        if stacked:
            stacked.gender_correct(genders)
    counts_all = stacked.counts if stacked else None
    totals = stacked.totals if stacked else None
    total_mask, bins_per_chr = get_mask(samples, counts_all, None, totals)
    g = np.array(genders)
    if genders.count("F") > 4:
        total_mask = total_mask & get_mask(samples[g == "F"], counts_all, np.flatnonzero(g == "F"), totals)[0]
    if genders.count("M") > 4 and not args.nipt:
        total_mask = total_mask & get_mask(samples[g == "M"], counts_all, np.flatnonzero(g == "M"), totals)[0]
    device = getattr(args, "device", 0)
    gpus = max(1, int(getattr(args, "gpus", 1) or 1))
    devices = [device + g for g in range(gpus)]
    parts = max(1, int(args.cpus))
    results = []
    timings["gender_model_and_masks"] = time.perf_counter() - t0

    writer = npz_io.AsyncNpzWriter(args.outfile)  # deflates the arrays of a finished pass while the next one runs

    import threading
    copier = [None, None]  # background thread of the previous pass, its exception

    def wait_copier():
        if copier[0] is not None:
            copier[0].join()
            copier[0] = None
        if copier[1] is not None:
            raise copier[1]

    def one_pass(sample_list, gender, nparts):
        t1 = time.perf_counter()
        # the pass's count matrix out of the stacked one: its chromosomes are a row prefix, its samples a column subset
        last = {"A": 22, "F": 23}.get(gender, 24)
        pass_counts = None
        if counts_all is not None and [max(len(s[str(c)]) for s in sample_list) for c in range(1, 25)] == list(bins_per_chr):
            rows_of_pass = int(sum(bins_per_chr[:last]))
            pass_counts = counts_all[:rows_of_pass] if gender == "A" else newref_tools.take_columns(counts_all, rows_of_pass, np.flatnonzero(g == gender))
        prep = newref_control.tool_newref_prep(sample_list, gender, total_mask, bins_per_chr, device, counts=pass_counts)
        t2 = time.perf_counter()
        wait_copier()  # the page-locked staging arrays are free again (the copy ran beside the preparation above)
        res = newref_control.tool_newref_main(prep, args.refsize, nparts, device, devices)
        results.append(res)

        def copy_out():
            # the GPU wrote into page-locked staging memory that the next pass reuses: move the arrays to ordinary
            # memory (NumPy releases the GIL for the copy) and start deflating them
            try:
                from concurrent.futures import ThreadPoolExecutor
                with ThreadPoolExecutor(4) as pool:  # first touch of 0.9 GB of fresh pages: a few threads, not one
                    for key in ("indexes", "distances", "null_ratios"):
                        src = res[key]
                        dst = np.empty(src.shape, src.dtype)
                        step = max(1, -(-len(src) // 8))
                        list(pool.map(lambda a: np.copyto(dst[a:a + step], src[a:a + step]), range(0, len(src), step)))
                        res[key] = dst
                        del src
                newref_control.writer_add_pass(writer, res, args.binsize)
            except Exception as e:  # re-raised on the main thread
                copier[1] = e

        copier[0] = threading.Thread(target=copy_out, name="wcx-copy-out")
        copier[0].start()
        timings["prep." + gender] = t2 - t1
        timings["get_reference." + gender] = time.perf_counter() - t2

    if len(genders) > 9:
        logging.info("Starting autosomal reference creation ...")
        logging.info("This might take a while ...")
        one_pass(list(samples), "A", parts)
    else:
        logging.critical("Provide at least 10 samples to enable the generation of a reference.")
        sys.exit()
    if genders.count("F") > 4:
        logging.info("Starting female gonosomal reference creation ...")
        one_pass(list(samples[g == "F"]), "F", 1)
    else:
        logging.warning("Provide at least 5 female samples to enable normalization of female gonosomes.")
    if not args.nipt:
        if genders.count("M") > 4:
            logging.info("Starting male gonosomal reference creation ...")
            one_pass(list(samples[g == "M"]), "M", 1)
        else:
            logging.warning("Provide at least 5 male samples to enable normalization of male gonosomes.")
    t0 = time.perf_counter()
    wait_copier()
    final_ref = newref_control.tool_newref_merge(args.outfile, results, args.binsize, args.nipt, trained_cutoff, writer)
    timings["write_reference"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    logging.info("Running QC on the newly created reference...")
    ref_qc.qc_reference(args.outfile, final_ref)  # reference main.py:134-135 (ref_qc.py:140); arrays still in memory
    timings["qc_reference"] = time.perf_counter() - t0
    logging.info("Stage wall-clock [s]: " + ", ".join("{} {:.2f}".format(k, v) for k, v in timings.items()))
    logging.info("Finished creating reference")
    return timings


# ---------------------------------------------------------------------------------------------
# predict
# ---------------------------------------------------------------------------------------------
get_post_processed_result = predict_control.get_post_processed_result
log_trans = predict_control.log_trans
apply_blacklist = predict_control.apply_blacklist


def tool_test(args):
    logging.info("Starting CNA prediction")
    if not args.bed and not args.plot:
        logging.critical("No output format selected. Select at least one of the supported output formats (--bed, --plot)")
        sys.exit()
    if args.zscore <= 0:
        logging.critical("Parameter --zscore should be a strictly positive number")
        sys.exit()
    if args.beta is not None and (args.beta <= 0 or args.beta > 1):
        logging.critical("Parameter --beta should be a strictly positive number lower than or equal to 1")
        sys.exit()
    if args.alpha <= 0 or args.alpha > 1:
        logging.critical("Parameter --alpha should be a strictly positive number lower than or equal to 1")
        sys.exit()
    if args.plot and not args.bed:
        logging.critical("--plot needs the reference's R plotter and is not part of the accelerated path; add --bed "
                         "(the tables are written here) or run the reference's `WisecondorX predict --plot`")
        sys.exit()
    logging.info("Importing data ...")
    # inflate the reference once (the reference re-inflates on every access)
    timings = {}
    t0 = time.perf_counter()
    _lib.prewarm_async(getattr(args, "device", 0))  # CUDA context creation runs beside the inflate of the reference
    ref_file = npz_io.load_npz(args.reference)  # all members inflated once, concurrently
    timings["load_reference"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    sample_file = np.load(args.infile, encoding="latin1", allow_pickle=True)
    engine = predict_tools.PredictEngine(getattr(args, "device", 0))
    # gender, both normalisations, assembly (reference main.py:242-271), log transform, blacklist, CBS + segment z-scores
    rem_input, results = predict_control.predict_batch(args, [sample_file["sample"].item()], [int(sample_file["binsize"].item())],
                                                       ref_file, engine, timings)[0]
    t0 = time.perf_counter()
    if args.bed:
        logging.info("Writing tables ...")
        predict_output.generate_output_tables(rem_input, results, engine)
    timings["write_tables"] = time.perf_counter() - t0
    logging.info("Stage wall-clock [s]: " + ", ".join("{} {:.2f}".format(k, v) for k, v in timings.items()))
    results["timings"] = timings
    if args.plot:
        logging.warning("--plot needs the reference's R plotter and is not part of the accelerated path; skipped")
    logging.info("Finished prediction")
    return results


def output_gender(args):
    ref_file = np.load(args.reference, encoding="latin1", allow_pickle=True)
    sample_file = np.load(args.infile, encoding="latin1", allow_pickle=True)
    print("male" if predict_gender(sample_file["sample"].item(), ref_file["trained_cutoff"]) == "M" else "female")


def tool_convert(args):
    logging.critical("`convert` (BAM/CRAM ingest through pysam) is not part of the accelerated path; "
                     "run the reference's `WisecondorX convert` -- its .npz output is read unchanged by newref / predict here")
    sys.exit()


def build_parser():
    """Same sub-commands, flags, types and defaults as the reference (main.py:312-488); `--device` is new."""
    parser = argparse.ArgumentParser(description="WisecondorX (B200-native numeric core)")
    parser.add_argument("--loglevel", type=str, default="INFO", choices=["info", "warning", "debug", "error", "critical"])
    sub = parser.add_subparsers()
    p = sub.add_parser("convert", description="Convert and filter aligned reads to .npz (reference only)")
    p.add_argument("infile", type=str); p.add_argument("outfile", type=str)
    p.add_argument("-r", "--reference", type=str); p.add_argument("--binsize", type=float, default=5e3)
    p.add_argument("--normdup", action="store_true")
    p.set_defaults(func=tool_convert)
    p = sub.add_parser("newref", description="Create a new reference using healthy reference samples")
    p.add_argument("infiles", type=str, nargs="+"); p.add_argument("outfile", type=str)
    p.add_argument("--nipt", action="store_true"); p.add_argument("--yfrac", type=float, default=None)
    p.add_argument("--plotyfrac", type=str, default=None); p.add_argument("--refsize", type=int, default=300)
    p.add_argument("--binsize", type=int, default=1e5); p.add_argument("--cpus", type=int, default=1)
    p.add_argument("--device", type=int, default=0, help="CUDA device (new)")
    p.add_argument("--gpus", type=int, default=1, help="number of GPUs (devices --device .. --device + N - 1) the target bins "
                                                        "are spread over (new; the reference spreads them over --cpus threads)")
    p.set_defaults(func=tool_newref)
    p = sub.add_parser("gender", description="Returns the gender of a .npz resulting from convert")
    p.add_argument("infile", type=str); p.add_argument("reference", type=str)
    p.set_defaults(func=output_gender)
    p = sub.add_parser("predict", description="Find copy number aberrations")
    p.add_argument("infile", type=str); p.add_argument("reference", type=str); p.add_argument("outid", type=str)
    p.add_argument("--minrefbins", type=int, default=150); p.add_argument("--maskrepeats", type=int, default=5)
    p.add_argument("--alpha", type=float, default=1e-4); p.add_argument("--zscore", type=float, default=5)
    p.add_argument("--beta", type=float, default=None); p.add_argument("--blacklist", type=str, default=None)
    p.add_argument("--gender", type=str, choices=["F", "M"]); p.add_argument("--ylim", type=str, default="def")
    p.add_argument("--bed", action="store_true"); p.add_argument("--plot", action="store_true")
    p.add_argument("--cairo", action="store_true"); p.add_argument("--add-plot-title", action="store_true")
    p.add_argument("--seed", type=int, default=None); p.add_argument("--regions", type=str, default=None)
    p.add_argument("--device", type=int, default=0, help="CUDA device (new)")
    p.set_defaults(func=tool_test)
    return parser


def main(argv=None):
    warnings.filterwarnings("ignore")
    parser = build_parser()
    args = parser.parse_args(argv)
    logging.basicConfig(format="[%(levelname)s - %(asctime)s]: %(message)s", datefmt="%Y-%m-%d %H:%M:%S",
                        level=getattr(logging, args.loglevel.upper(), None))
    logging.debug("args are: {}".format(args))
    if not hasattr(args, "func"):
        parser.print_help()
        return
    args.func(args)


if __name__ == "__main__":
    main()
