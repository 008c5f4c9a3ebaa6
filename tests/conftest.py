"""pytest configuration: registers the ``gpu`` marker and puts the package on sys.path.

``-m "not gpu"`` runs here (no GPU): oracle vs golden fixtures, host logic, C-ABI symbol checks.
``-m gpu`` runs on a B200: the parity tests proper, through the C-ABI.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "relevance-factorizationmachine_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_csr(g, prefix):
    from scipy.sparse import csr_matrix
    shape = tuple(int(v) for v in g[prefix + "_shape"])
    return csr_matrix((g[prefix + "_data"], g[prefix + "_indices"], g[prefix + "_indptr"]), shape=shape)


def golden_frame(g):
    return {c: g["frame_" + c] for c in ("user", "item", "label", "pscore", "ones_pscore")}


@pytest.fixture(scope="session")
def golden():
    return load_golden
