#!/usr/bin/env python3
"""Generate the Boys-function interpolation table used by the CUDA kernels.

F_m(T) = int_0^1 u^(2m) exp(-T u^2) du is interpolated on [0, TMAX) with
NINT = TMAX/DELTA intervals of width DELTA = 1/7; inside interval iv the value is
a degree-7 polynomial in xd = T/DELTA - iv - 1/2 (xd in [-1/2, 1/2]) evaluated by
Horner's rule.  These are the same interpolation parameters as the reference's
FmEval_Chebyshev7 (include/libint2/boys.h:253-482; order 7, Tmax 117, delta 1/7,
boys_cheb7_v2.h:24-28) so that the device evaluator follows the same branches;
the coefficients themselves are computed here from scratch with mpmath:
degree-7 interpolation at the 8 Chebyshev nodes of each interval, converted to
monomial coefficients in xd, rounded once to double.

Output: libint_b200/data/boys_cheb7_m{MMAX}.bin -- little-endian float64,
shape [NINT][MMAX+1][8], preceded by no header (shape is implied by file name).
"""
import os
import sys
import numpy as np
import mpmath as mp

ORDER = 7
TMAX = 117
NINT = TMAX * 7  # 819
MMAX = int(sys.argv[1]) if len(sys.argv) > 1 else 24

mp.mp.dps = 50


def boys_all(T, mmax):
    """F_m(T) for m = 0..mmax: F_mmax from 1F1, then downward recursion."""
    T = mp.mpf(T)
    out = [None] * (mmax + 1)
    out[mmax] = mp.hyp1f1(mmax + mp.mpf(1) / 2, mmax + mp.mpf(3) / 2, -T) / (2 * mmax + 1)
    eT = mp.exp(-T)
    for m in range(mmax - 1, -1, -1):
        out[m] = (2 * T * out[m + 1] + eT) / (2 * m + 1)
    return out


def cheb_to_mono(a):
    """coefficients of sum_k a_k T_k(t) as a polynomial in t (mp numbers)."""
    n = len(a)
    Tk = [[mp.mpf(1)], [mp.mpf(0), mp.mpf(1)]]
    for k in range(2, n):
        prev, prev2 = Tk[k - 1], Tk[k - 2]
        cur = [mp.mpf(0)] * (k + 1)
        for i, c in enumerate(prev):
            cur[i + 1] += 2 * c
        for i, c in enumerate(prev2):
            cur[i] -= c
        Tk.append(cur)
    mono = [mp.mpf(0)] * n
    for k in range(n):
        for i, c in enumerate(Tk[k]):
            mono[i] += a[k] * c
    return mono


def main():
    n = ORDER + 1
    nodes = [mp.cos(mp.pi * (2 * j + 1) / (2 * n)) for j in range(n)]  # t in (-1,1)
    table = np.zeros((NINT, MMAX + 1, n))
    delta = mp.mpf(1) / 7
    for iv in range(NINT):
        x0 = (iv + mp.mpf(1) / 2) * delta
        vals = [boys_all(x0 + t * delta / 2, MMAX) for t in nodes]  # x = x0 + xd*delta, xd = t/2
        for m in range(MMAX + 1):
            f = [vals[j][m] for j in range(n)]
            a = []
            for k in range(n):
                s = sum(f[j] * mp.cos(mp.pi * k * (2 * j + 1) / (2 * n)) for j in range(n))
                a.append(s * (1 if k == 0 else 2) / n)
            mono_t = cheb_to_mono(a)  # polynomial in t = 2*xd
            for k in range(n):
                table[iv, m, k] = float(mono_t[k] * mp.mpf(2) ** k)
        if iv % 100 == 0:
            print("interval", iv, flush=True)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data",
                       "boys_cheb7_m%d.bin" % MMAX)
    table.astype("<f8").tofile(out)
    print("wrote", os.path.normpath(out), table.shape)


if __name__ == "__main__":
    main()
