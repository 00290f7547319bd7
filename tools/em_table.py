#!/usr/bin/env python3
"""Coefficients of the Euler-Maclaurin form of the integrated Gaussian used by the narrow engine (od_narrow.cu).

    omega Phi(z_j) = dE S_j - g_j [dE/2 + w z_j W(z_j^2)],   z W(z^2) = - sum_{k=1..NT} B_2k/(2k)! del^2k He_{2k-1}(z)

(S_j = inclusive sum of the record's Gaussian values g_i on the grid, del = dE/w, He_n the probabilists' Hermite
polynomials, B_2k the Bernoulli numbers).  Prints the table T[k][m] = - B_2k/(2k)! * [z^(2m+1)] He_{2k-1}(z) and, for
every NT, the first omitted term's bound  |B_{2NT+2}/(2NT+2)!| del^(2NT+2) max_z |He_{2NT+1}(z) phi(z)|  as a function
of del, from which the del thresholds of the kernel are read (target: <= 3e-12 of omega)."""
from fractions import Fraction
from math import comb, factorial

import numpy as np


def bernoulli(n):
    B = [Fraction(0)] * (n + 1)
    B[0] = Fraction(1)
    for m in range(1, n + 1):
        B[m] = -sum(comb(m + 1, k) * B[k] for k in range(m)) / (m + 1)
    return B


def hermite_e(n):
    """coefficients c[p] of z^p in He_n"""
    h0, h1 = [Fraction(1)], [Fraction(0), Fraction(1)]
    if n == 0:
        return h0
    for k in range(1, n):
        nxt = [Fraction(0)] + h1              # z * He_k
        for p, c in enumerate(h0):
            nxt[p] -= k * c                   # - k He_{k-1}
        h0, h1 = h1, nxt
    return h1


def main():
    NTMAX = 10
    B = bernoulli(2 * NTMAX + 2)
    z = np.linspace(0, 12, 240001)
    phi = np.exp(-0.5 * z * z) / np.sqrt(2 * np.pi)
    print("// T[k-1][m] = -B_2k/(2k)! * coefficient of z^(2m+1) in He_{2k-1}   (tools/em_table.py)")
    for k in range(1, NTMAX + 1):
        he = hermite_e(2 * k - 1)
        row = [float(-B[2 * k] / factorial(2 * k) * he[2 * m + 1]) if 2 * m + 1 < len(he) else 0.0 for m in range(NTMAX)]
        print("  {" + ", ".join(f"{v:.17e}" for v in row) + "},")
    print()
    for NT in range(3, NTMAX + 1):
        k = NT + 1
        he = hermite_e(2 * k - 1)
        poly = sum(float(c) * z ** p for p, c in enumerate(he))
        mx = np.abs(poly * phi).max()
        coef = abs(float(B[2 * k] / factorial(2 * k))) * mx
        dmax = (3e-12 / coef) ** (1.0 / (2 * k))
        print(f"NT={NT}: first omitted term <= {coef:.3e} * del^{2 * k};  <= 3e-12 for del <= {dmax:.4f}  (windows >= {16.4 / dmax + 1:.1f} bins)")


if __name__ == "__main__":
    main()
