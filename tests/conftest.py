import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on a B200)")


def load_npz(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.fixture(scope="session")
def odo():
    """The CPU oracle (test infrastructure only)."""
    from oracle import odo as _odo
    _odo.build()
    _odo.lib()
    return _odo


def rel_to_peak(a, b):
    """max |a-b| relative to the peak of the reference column b."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    peak = np.abs(b).max()
    if peak == 0.0:
        return float(np.abs(a - b).max())
    return float(np.abs(a - b).max() / peak)
