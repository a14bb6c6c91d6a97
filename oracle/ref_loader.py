"""TEST INFRASTRUCTURE ONLY -- loader for the *live* reference (WisecondorX v1.2.10).

Puts ``/root/reference/src`` on ``sys.path`` with the shims SURVEY.md section 8c lists so the
reference's own NumPy functions can be called to (a) pin the restatements in
``oracle/np_oracle.py`` / ``oracle/wcx_oracle.c`` and (b) generate the golden vectors under
``tests/golden/`` (``tests/golden/make_golden.py``).

``/root/reference`` exists only in the build container: nothing on the GPU box may import
this module (``available()`` is False there and the pin tests skip).  Nothing in the product
package ``wisecondorx_b200`` imports anything under ``oracle/``.
"""
from __future__ import annotations

import os
import sys
import types

REF_SRC = "/root/reference/src"


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "wisecondorx"))


_loaded = None


def load():
    """Returns a namespace with the reference modules (newref_tools, newref_control,
    predict_tools, predict_control, overall_tools, main, predict_output, ref_qc)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("live reference not present (expected only in the build container)")
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    # pysam is absent; convert_tools imports it at module scope (convert_tools.py:6)
    sys.modules.setdefault("pysam", types.ModuleType("pysam"))
    import numpy as np
    from wisecondorx import (newref_tools, newref_control, predict_tools, predict_control,
                             overall_tools)
    from sklearn.decomposition import PCA

    # sklearn >= 1.5: PCA.transform touches explained_variance_, which project_pc never sets
    # (predict_tools.py:57-59).  whiten=False so the value is unused.
    class _ShimPCA(PCA):
        explained_variance_ = np.ones(1)

    predict_tools.PCA = _ShimPCA
    from wisecondorx import main, predict_output, ref_qc
    main.qc_reference = ref_qc.qc_reference  # main.py:135 calls it without importing it (NameError)
    ns = types.SimpleNamespace(newref_tools=newref_tools, newref_control=newref_control,
                               predict_tools=predict_tools, predict_control=predict_control,
                               overall_tools=overall_tools, main=main, predict_output=predict_output, ref_qc=ref_qc)
    _loaded = ns
    return ns
