import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session", autouse=True)
def _library_present():
    """The host-logic tests call the library's host-side entry points (no GPU needed): build it when the tree has no
    .so yet (a fresh clone; nvcc cross-compiles without a GPU).  A .so that is there is left alone -- on the GPU box the
    one that travelled with the snapshot is the one under test."""
    from wisecondorx_b200 import _lib, build
    if not os.path.exists(_lib.SO_PATH):
        build.build()
