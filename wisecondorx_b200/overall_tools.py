"""Host helpers shared by newref and predict (mirror of the reference's overall_tools.py; cheap,
O(bins) work that stays on the host).  get_z_score lives in predict_tools (GPU)."""
from __future__ import annotations

import logging
import sys

import numpy as np


def scale_sample(sample, from_size, to_size):
    """Re-bins a sample to a coarser bin size (reference overall_tools.py:19-40), vectorised with
    np.add.reduceat.  Same error convention: log + sys.exit() (status 0)."""
    if not to_size or from_size == to_size:
        return sample
    if to_size == 0 or from_size == 0 or to_size < from_size or to_size % from_size > 0:
        logging.critical("Impossible binsize scaling requested: {} to {}".format(int(from_size), int(to_size)))
        sys.exit()
    scale = int(to_size // from_size)
    out = {}
    for name, data in sample.items():
        data = np.asarray(data)
        new_len = int(np.ceil(len(data) / float(scale)))
        if new_len == 0:
            continue  # the reference assigns inside its per-bin loop (:38-39): an empty chromosome leaves no key
        out[name] = np.add.reduceat(data, np.arange(0, len(data), scale)).astype(np.int32)
    return out


def gender_correct(sample, gender):
    """Levels the gonosomal read counts of males with the autosomes (reference :48-53)."""
    if gender == "M":
        sample["23"] = sample["23"] * 2
        sample["24"] = sample["24"] * 2
    return sample


def get_median_segment_variance(results_c, results_r):
    """Median over segments of the variance of their non-zero ratios (reference :127-135)."""
    variances = []
    for seg in results_c:
        r = np.asarray(results_r[seg[0]][int(seg[1]):int(seg[2])], dtype=float)
        r = r[r != 0]
        if len(r):
            variances.append(np.var(r))
    return np.median(variances)


def get_cpa(results_c, binsize):
    """Copy number profile abnormality score (reference :143-148)."""
    x = 0.0
    for seg in results_c:
        x += (seg[2] - seg[1] + 1) * binsize * abs(seg[3])
    return x / len(results_c) * (10 ** -8)
