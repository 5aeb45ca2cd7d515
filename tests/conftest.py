import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_pileup(g, prefix=""):
    from secedo_b200.pileup import Pileup
    return Pileup(g[prefix + "chr_ptr"], g[prefix + "row_ptr"], g[prefix + "position"], g[prefix + "read_id"],
                  g[prefix + "gid_base"])


def assert_matrix_close(M, ref, tol=1e-6):
    """|M - ref| <= tol * max|ref| (absolute, scaled — ADD_MIN makes the smallest
    entry exactly 0, so an element-relative bound is meaningless). NaNs (0 * inf of SCALE_MAX_1 on
    an all-zero matrix) must coincide."""
    M, ref = np.asarray(M), np.asarray(ref)
    assert M.shape == ref.shape
    nan_m, nan_r = np.isnan(M), np.isnan(ref)
    assert np.array_equal(nan_m, nan_r)
    scale = np.nanmax(np.abs(ref)) if (~nan_r).any() else 0.0
    diff = np.nanmax(np.abs(M - ref)) if (~nan_r).any() else 0.0
    assert diff <= tol * max(scale, 1e-300), f"max|diff|={diff} scale={scale}"


@pytest.fixture(scope="session")
def gpu_ctx():
    from secedo_b200.api import Context
    return Context(0)
