import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def golden_names(prefix):
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith(prefix) and f.endswith(".npz"))


def rel_err(a, b):
    """max |a-b| / max(|b|, tiny): the per-iterate relative error of the north star."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def assert_bitwise(a, b, what=""):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    bad = a != b  # -0.0 == +0.0, as intended
    if bad.any():
        idx = np.argwhere(bad)
        raise AssertionError("%s: %d/%d values differ, first at %s: %r vs %r (max abs diff %.3g)" % (
            what, bad.sum(), a.size, idx[0], a[tuple(idx[0])], b[tuple(idx[0])],
            np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))))


@pytest.fixture(scope="session")
def oracle():
    import oracle as orc
    orc.build()
    return orc
