"""Pins the CBS restatement (oracle/cbs_oracle.py) -- and, with a GPU, the CUDA path -- to the REAL DNAcopy when
tests/golden/cbs_dnacopy.json exists.  The file is produced by tools/make_cbs_golden.R at a site that has R + DNAcopy
(the build image has neither: SURVEY.md 8c); without it these tests skip and CBS parity stays "unpinned"."""
import json
import os

import numpy as np
import pytest

from oracle import cbs_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cbs_dnacopy.json")
pytestmark = pytest.mark.skipif(not os.path.exists(GOLD), reason="no DNAcopy fixtures (run tools/make_cbs_golden.R where R is installed)")


def _cases():
    return json.load(open(GOLD))["cases"]


def _ends(case):
    """DNAcopy's loc.end are 1-based positions in the full vector (NA bins carry 0 here, as CBS.R:41 reads them)."""
    return [int(e) for e in case["loc_end"]]


def _oracle_ends(case):
    y = np.asarray(case["y"], dtype=float)
    w = np.asarray(case["w"], dtype=float)
    keep = np.flatnonzero(y != 0)
    ends = cbs_oracle.segment_chromosome(y[keep], w[keep], alpha=case["alpha"], seed=case["seed"])
    return [int(keep[e - 1]) + 1 for e in ends]


def test_oracle_breakpoints_match_dnacopy():
    bad = {}
    for name, case in _cases().items():
        got, want = _oracle_ends(case), _ends(case)
        if got != want:
            bad[name] = (got, want)
    assert not bad, bad


@pytest.mark.gpu
def test_gpu_breakpoints_match_dnacopy():
    from wisecondorx_b200 import cbs
    series, keeps = [], []
    cases = _cases()
    for case in cases.values():
        y = np.asarray(case["y"], dtype=float)
        w = np.asarray(case["w"], dtype=float)
        keep = np.flatnonzero(y != 0)
        keeps.append(keep)
        series.append((y[keep], w[keep]))
    ends = cbs.segment_series(series, [0] * len(series), alpha=1e-4, seed=1)
    for (name, case), keep, e in zip(cases.items(), keeps, ends):
        assert [int(keep[i - 1]) + 1 for i in e] == _ends(case), name
