"""Numpy model of the kernel's index algebra for a general plan N = R*C (plan 400: R=C=20; plan 512: R=32, C=16):
column DFTs of length R, Z exchange, per-worker twiddle + conjugate/rotation trick, row DFTs of length C, untangle of
the two packed real frames, power-row map.  Development aid; validates the tables the CUDA kernels hard-code."""
import sys
import numpy as np


def model_pair(fa, fb, R, C):
    N = R * C
    W = R // 2                                   # workers per FFT
    z = fa + 1j * fb
    Z = np.zeros((R, C), dtype=complex)          # [row k1][n2]
    for c in range(C):
        y = np.fft.fft(z[C * np.arange(R) + c])  # Y[c][k1]
        y[R // 2] *= np.exp(+2j * np.pi * c / (2 * C))      # row R/2 pre-rotated by W_{2C}^{-c}
        Z[:, c] = y
    n2 = np.arange(C)
    PA = np.full(N // 2 + 1, np.nan)
    PB = np.full(N // 2 + 1, np.nan)
    rows = {}
    for t in range(W):
        tw = np.exp(-2j * np.pi * t * n2 / N)
        r1, r2 = t, (R - t) if t else R // 2
        X = np.fft.fft(Z[r1] * tw)
        D = np.fft.fft(Z[r2] * np.conj(tw))
        Y = np.roll(D, -1)
        lo_base, hi_base = (R // 2, R) if t == 0 else (t, R - t)
        for j in range(C):
            if t == 0:
                u = Y[j] if j < C // 2 else X[j]
                v = X[C - j] if j >= C // 2 else Y[C - 1 - j]
            else:
                u, v = X[j], Y[C - 1 - j]
            pa = ((u.real + v.real) ** 2 + (u.imag - v.imag) ** 2) * 0.25
            pb = ((u.imag + v.imag) ** 2 + (u.real - v.real) ** 2) * 0.25
            b = lo_base + R * j if j < C // 2 else hi_base + R * (C - 1 - j)
            row = W * j + t
            assert row not in rows
            rows[row] = b
            PA[b], PB[b] = pa, pb
    return PA, PB, rows


def row_of_bin(b, R, C):
    W = R // 2
    rr, q = b % R, b // R
    if rr == 0:
        t, j = 0, C - q
    elif rr == R // 2:
        t, j = 0, q
    elif rr < R // 2:
        t, j = rr, q
    else:
        t, j = R - rr, C - 1 - q
    return W * j + t


if __name__ == "__main__":
    for R, C in ((20, 20), (32, 16)):
        N = R * C
        rng = np.random.default_rng(0)
        fa, fb = rng.standard_normal(N), rng.standard_normal(N)
        PA, PB, rows = model_pair(fa, fb, R, C)
        ra, rb = np.abs(np.fft.fft(fa)[:N // 2 + 1]) ** 2, np.abs(np.fft.fft(fb)[:N // 2 + 1]) ** 2
        m = ~np.isnan(PA)
        bins = sorted(set(rows.values()))
        ok_rows = all(row_of_bin(b, R, C) == r for r, b in rows.items())
        print(N, "bins", bins[0], bins[-1], "missing", sorted(set(range(N // 2 + 1)) - set(bins)),
              "errA", np.abs(PA[m] - ra[m]).max() / ra.max(), "errB", np.abs(PB[m] - rb[m]).max() / rb.max(),
              "row_of_bin ok", ok_rows)
