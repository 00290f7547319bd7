"""The Euler-Maclaurin end-correction filter of the narrow engine (optados_b200/csrc/od_narrow.cu, c_em_fir; designed by
tools/em_fir.py): the taps compiled into the library must turn the prefix sum of a sampled Gaussian into its integral,

    F(E_j) = dE sum_{n<=j} f_n - dE [ f_j / 2 + sum_n a_n (f_{j+n} - f_{j-n}) ],

to ~1e-15 of the Gaussian's weight for every width the kernel sends through it (dE / w <= EM_DELMAX) and every offset of
the centre from the grid.  CPU only: it reads the constants from the source."""
import os
import re
from math import erf, pi, sqrt

import numpy as np

SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "optados_b200", "csrc", "od_narrow.cu")


def _constants():
    s = open(SRC).read()
    delmax = float(re.search(r"constexpr double EM_DELMAX = ([0-9.]+);", s).group(1))
    ntaps = int(re.search(r"constexpr int EM_NTAPS = (\d+);", s).group(1))
    body = re.search(r"c_em_fir\[EM_NTAPS\] = \{(.*?)\};", s, re.S).group(1)
    taps = [float(x) for x in re.findall(r"[-+]?\d\.\d+e[-+]\d+", body)]
    assert len(taps) == ntaps
    return delmax, np.array(taps)


def _worst(taps, dl, offsets):
    n = len(taps)
    w = 1.0 / dl
    half = int(8.3 * w) + 2
    j = np.arange(-half - n - 2, half + n + 3)
    worst = 0.0
    for s in offsets:
        z = (j - s) / w
        f = np.where(np.abs(z) <= 8.0 + 1.0 / w, np.exp(-0.5 * z * z) / (w * sqrt(2 * pi)), 0.0)   # deposited over >= +-8 w, as the kernel does
        fir = np.zeros_like(f)
        for k, t in enumerate(taps, 1):
            fir[k:-k] += t * (f[2 * k:] - f[:-2 * k])
        F = np.cumsum(f) - (0.5 * f + fir)
        exact = np.array([0.5 * (1.0 + erf(v / sqrt(2.0))) for v in z])
        worst = max(worst, np.abs(F - exact)[n + 2:-n - 2].max())
    return worst


def test_filter_integrates_sampled_gaussians():
    delmax, taps = _constants()
    assert 0.3 <= delmax <= 0.45
    offsets = (0.0, 0.13, 0.37, 0.5, 0.77)
    for dl in (delmax, 0.97 * delmax, 0.9 * delmax, 0.3, 0.2, 0.1, 0.05, 0.02):
        # (the +-8 w deposit leaves tails of exp(-32) = 1.3e-14 of the peak; the filter itself is good to 1.3e-15)
        assert _worst(taps, dl, offsets) < 2e-14, dl


def test_filter_fails_outside_its_band():
    """the guard of the kernel (dE / w <= EM_DELMAX) is needed: at dE / w = 0.6 the sampled Gaussian aliases"""
    _, taps = _constants()
    assert _worst(taps, 0.6, (0.0, 0.37)) > 1e-9
