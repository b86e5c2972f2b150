import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def demo():
    """Golden fixture: reference demo molecules + oracle outputs (tests/golden/make_golden.py)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "demo.npz"))


def mbtypes_from_array(mb):
    return [[int(v) for v in row if v >= 0] for row in mb]


def rel_err(got, ref, floor=None):
    """max |got-ref| / max(|ref|, floor); floor defaults to 1e-10 * max|ref| (hard part 9)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if floor is None:
        floor = 1e-10 * max(np.max(np.abs(ref)), 1e-300)
    return float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), floor))) if ref.size else 0.0


def kernel_rel_err(got, ref):
    """Per-element RELATIVE error of a kernel matrix (north_star: "1e-10 relative per kernel element").  No floor
    tied to the largest element: the only floor is where exp itself leaves the normal double range (|ref| < 1e-290:
    the reference's libm returns denormals down to exp(-745), the device flushes results below 2^-1022 to zero), there
    the difference is taken absolutely."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-290))) if ref.size else 0.0
