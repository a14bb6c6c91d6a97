"""Output tables of `predict --bed` (formats of the reference's predict_output.py:59-261), written
with vectorised host code.  Plotting (`--plot`) needs R and is out of scope."""
from __future__ import annotations

import numpy as np

from . import overall_tools, predict_tools


def _chr_name(c):
    return {22: "X", 23: "Y"}.get(c, str(c + 1))


def _fmt(v):
    return "nan" if (not isinstance(v, str) and v == 0) else str(v)


def generate_output_tables(rem_input, results, engine=None):
    _generate_bins_bed(rem_input, results)
    _generate_segments_and_aberrations_bed(rem_input, results)
    _generate_chr_statistics_file(rem_input, results, engine)
    if getattr(rem_input["args"], "regions", None) is not None:
        _generate_regions_bed(rem_input, results)


def _generate_regions_bed(rem_input, results):
    """Weighted mean ratio / z-score of user-given regions (reference predict_output.py:86-137).  The reference maps
    chrX / chrY to indexes 21 / 22 and then overrides the result with int(name) (a ValueError for X / Y, SURVEY.md
    A.8); here X / Y address chromosomes 23 / 24 of the results, everything else is as in the reference."""
    import re
    binsize = rem_input["binsize"]
    with open("{}_regions.bed".format(rem_input["args"].outid), "w") as fh, open(rem_input["args"].regions) as rh:
        fh.write("chr\tstart\tend\tname\tratio\tzscore\n")
        for region in (line.strip().split("\t") for line in rh if line.strip() != ""):
            assert len(region) >= 4, "Regions file must have at least 4 columns: chr, start, end, name"
            chr_name, start, end, name = region[:4]
            short = re.sub("chr", "", chr_name)
            c = {"X": 22, "Y": 23}.get(short)
            c = int(short) - 1 if c is None else c
            start_bin, end_bin = int(start) // binsize, int(end) // binsize
            if c >= len(results["results_r"]):
                fh.write("Skipping invalid region: {}\n".format("\t".join(region)))
                continue
            end_bin = min(end_bin, int(rem_input["bins_per_chr"][c]) - 1)
            if start_bin < 0 or end_bin < 0 or start_bin > end_bin:
                fh.write("Skipping invalid region: {}\n".format("\t".join(region)))
                continue
            r = np.asarray(results["results_r"][c][start_bin:end_bin + 1], dtype=float)
            w = np.asarray(results["results_w"][c][start_bin:end_bin + 1], dtype=float)
            z = np.asarray(results["results_z"][c][start_bin:end_bin + 1], dtype=float)
            if len(r) == 0:
                fh.write("Skipping region with no bins: {}\n".format("\t".join(region)))
                continue
            fh.write("\t".join(str(x) for x in [chr_name, start, end, name, _fmt(np.ma.average(r, weights=w)),
                                                _fmt(np.ma.average(z, weights=w))]) + "\n")


def _format_bins(name, binsize, r, z):
    import ctypes
    from . import _lib
    r, z = np.ascontiguousarray(r), np.ascontiguousarray(z)
    out = np.empty(len(r) * (2 * len(name) + 138) + 16, dtype=np.uint8)
    n = ctypes.c_int64()
    _lib.check(_lib.load().wcx_host_format_bins(name.encode("ascii"), binsize, ctypes.c_void_p(r.ctypes.data), ctypes.c_void_p(z.ctypes.data),
                                                len(r), ctypes.c_void_p(out.ctypes.data), out.nbytes, ctypes.byref(n)))
    return out[:n.value].tobytes().decode("ascii")


def _generate_bins_bed(rem_input, results):
    binsize = rem_input["binsize"]
    with open("{}_bins.bed".format(rem_input["args"].outid), "w") as fh:
        fh.write("chr\tstart\tend\tid\tratio\tzscore\n")
        for c in range(len(results["results_r"])):
            name = _chr_name(c)
            r, z = results["results_r"][c], results["results_z"][c]
            if (isinstance(r, np.ndarray) and isinstance(z, np.ndarray) and r.dtype == np.float64 and z.dtype == np.float64
                    and r.ndim == 1 and z.shape == r.shape and isinstance(binsize, (int, np.integer))):
                # 200 k lines at 15 kb: formatted by the library on the host (wcx_host_format_bins, csrc/host_tables.cu:
                # Python's repr of every value = str of the NumPy scalar the reference prints)
                fh.write(_format_bins(name, int(binsize), r, z))
                continue
            lines = []
            for i in range(len(r)):
                s, e = i * binsize + 1, (i + 1) * binsize
                lines.append("{}\t{}\t{}\t{}:{}-{}\t{}\t{}\n".format(name, s, e, name, s, e, _fmt(r[i]), _fmt(z[i])))
            fh.write("".join(lines))


def _aberration_cutoff(beta, ploidy):
    return np.log2((ploidy - (beta / 2)) / ploidy), np.log2((ploidy + (beta / 2)) / ploidy)


def _generate_segments_and_aberrations_bed(rem_input, results):
    args = rem_input["args"]
    with open("{}_segments.bed".format(args.outid), "w") as seg_fh, open("{}_aberrations.bed".format(args.outid), "w") as ab_fh:
        seg_fh.write("chr\tstart\tend\tratio\tzscore\n")
        ab_fh.write("chr\tstart\tend\tratio\tzscore\ttype\n")
        for seg in results["results_c"]:
            name = _chr_name(seg[0])
            row = "\t".join(str(x) for x in [name, int(seg[1] * rem_input["binsize"] + 1), int(seg[2] * rem_input["binsize"]), seg[4], seg[3]])
            seg_fh.write(row + "\n")
            ploidy = 1 if (name in ("X", "Y") and rem_input["ref_gender"] == "M") else 2
            if args.beta is not None:
                lo, hi = _aberration_cutoff(args.beta, ploidy)
                if float(seg[4]) > hi:
                    ab_fh.write(row + "\tgain\n")
                elif float(seg[4]) < lo:
                    ab_fh.write(row + "\tloss\n")
            elif isinstance(seg[3], str):
                continue
            elif float(seg[3]) > args.zscore:
                ab_fh.write(row + "\tgain\n")
            elif float(seg[3]) < -args.zscore:
                ab_fh.write(row + "\tloss\n")


def _generate_chr_statistics_file(rem_input, results, engine=None):
    nchr = len(results["results_r"])
    means, medians = [], []
    for c in range(nchr):
        r = np.asarray(results["results_r"][c], dtype=float)
        w = np.asarray(results["results_w"][c], dtype=float)
        means.append(np.ma.average(r, weights=w))
        nz = r[r != 0]
        medians.append(np.median(nz) if len(nz) else float("nan"))
    results_c_chr = [[c, 0, int(rem_input["bins_per_chr"][c]) - 1, means[c]] for c in range(nchr)]
    msv = round(overall_tools.get_median_segment_variance(results["results_c"], results["results_r"]), 5)
    cpa = round(overall_tools.get_cpa(results["results_c"], rem_input["binsize"]), 5)
    chr_z = predict_tools.get_z_score(results_c_chr, results, engine)
    with open("{}_statistics.txt".format(rem_input["args"].outid), "w") as fh:
        fh.write("chr\tratio.mean\tratio.median\tzscore\n")
        for c in range(nchr):
            fh.write("\t".join(str(x) for x in [_chr_name(c), means[c], medians[c], chr_z[c]]) + "\n")
        fh.write("Gender based on --yfrac (or manually overridden by --gender): {}\n".format(rem_input["gender"]))
        fh.write("Number of reads: {}\n".format(rem_input["n_reads"]))
        fh.write("Standard deviation of the ratios per chromosome: {}\n".format(round(float(np.nanstd(means)), 5)))
        fh.write("Median segment variance per bin (doi: 10.1093/nar/gky1263): {}\n".format(msv))
        fh.write("Copy number profile abnormality (CPA) score (doi: 10.1186/s13073-020-00735-4): {}\n".format(cpa))
