#!/usr/bin/env python3
"""Design of the end-correction filter of the integrated DOS (od_narrow.cu, Euler-Maclaurin groups).

For f sampled with spacing h,   h sum_{n<=j} f_n = F(E_j) + h [ f_j / 2 + sum_k B_2k h^(2k-1)/(2k)! f^(2k-1)(E_j) ]   (Euler-Maclaurin);
on f = exp(i k E) the bracket is  1/2 + i phi(theta),  phi(theta) = 1/theta - cot(theta/2)/2,  theta = k h.  A sum of Gaussians
of width w >= h / del_max sampled on the grid has the spectrum exp(-(theta/del)^2/2): for del_max = 0.3 it is below 1e-24 at the
Nyquist frequency, so the derivative series can be applied to the SUMMED spectrum as one antisymmetric FIR filter,
    sum_k ... = sum_{n=1..N} a_n (f_{j+n} - f_{j-n}),      2 sum_n a_n sin(n theta) = phi(theta)  on the band the data occupy,
instead of per (record, bin).  The taps minimise the error weighted with the widest spectrum (least squares, 60 digits).
usage: tools/em_fir.py [del_max] [N]   -> prints the taps and the worst error over widths and grid offsets"""
import sys
import mpmath as mp

mp.mp.dps = 60
dmax = mp.mpf(sys.argv[1]) if len(sys.argv) > 1 else mp.mpf("0.3")
N = int(sys.argv[2]) if len(sys.argv) > 2 else 32


def phi(t):
    return 1 / t - mp.cot(t / 2) / 2 if t != 0 else mp.mpf(0)


# Gauss-Legendre nodes on [0, pi]
M = 400
xs, ws = [], []
for lo, hi in [(mp.pi * i / 8, mp.pi * (i + 1) / 8) for i in range(8)]:
    for x, w in mp.calculus.quadrature.GaussLegendre(mp.mp).calc_nodes(6, mp.mp.prec):
        xs.append((hi - lo) / 2 * x + (hi + lo) / 2)
        ws.append((hi - lo) / 2 * w)
W2 = [mp.exp(-(x / dmax) ** 2) + mp.mpf(10) ** -40 for x in xs]  # weight squared: spectrum^2 (+ floor)
S = [[2 * mp.sin(n * x) for x in xs] for n in range(1, N + 1)]
P = [phi(x) for x in xs]
G = mp.matrix(N, N)
b = mp.matrix(N, 1)
for m in range(N):
    for n in range(m, N):
        G[m, n] = G[n, m] = mp.fsum(w * q * sm * sn for w, q, sm, sn in zip(ws, W2, S[m], S[n]))
    b[m] = mp.fsum(w * q * sm * p for w, q, sm, p in zip(ws, W2, S[m], P))
a = mp.lu_solve(G, b)
taps = [float(a[i]) for i in range(N)]
print("del_max", dmax, "N", N)
print("taps = {" + ", ".join(f"{t:.17e}" for t in taps) + "}")
print("sum|a|", sum(abs(t) for t in taps))

# check: one Gaussian of unit weight, width w = h/del, centre at grid offset s; F_j exact vs h*prefix - h*(f/2 + fir)
import numpy as np
from math import erf, sqrt, pi, exp
worst = 0.0
for dl in [float(dmax), 0.97 * float(dmax), 0.9 * float(dmax), 0.3, 0.25, 0.2, 0.15, 0.1, 0.05, 0.02]:
    if dl > float(dmax) + 1e-12:
        continue
    for s in [0.0, 0.13, 0.37, 0.5, 0.77]:
        w = 1.0 / dl
        half = int(8.3 * w) + 2
        j = np.arange(-half - N - 2, half + N + 3)
        z = (j - s) / w
        f = np.where(np.abs(z) <= 8.2024 + 1.0 / w, np.exp(-0.5 * z * z) / (w * sqrt(2 * pi)), 0.0)  # truncated like the window
        pre = np.cumsum(f)
        fir = np.zeros_like(f)
        for n, t in enumerate(taps, 1):
            fir[n:-n] += t * (f[2 * n:] - f[:-2 * n])
        F = pre - (0.5 * f + fir)
        ex = np.array([0.5 * (1 + erf(v / sqrt(2))) for v in z])
        err = np.abs(F - ex)[N + 2:-N - 2].max()
        worst = max(worst, err)
        print(f"del {dl:5.2f} offset {s:4.2f}  max |F - Phi| = {err:.2e}")
print("worst", worst)
