"""(The kernel now stores the powers in natural bin order, `b` below, not at `row = 10*j + t`; the (t, j) -> bin map and the
rest of the algebra are unchanged.)
Numpy model of the plan-400 kernel's index algebra (PFA DFT-20, 20x20 Cooley-Tukey, Z slot map,
two-frames-per-complex-FFT untangle, P row map).  Development aid: validates the tables/permutations the
CUDA kernel hard-codes.  Not used by the product or the tests' oracle."""
import numpy as np

N, R, C, G = 400, 20, 20, 10


def dft20_pfa(x):
    """20-point DFT as 4x5 Good-Thomas: n=(5a+4b)%20, k=(5ka+16kb)%20, no twiddles."""
    t = np.zeros((4, 5), dtype=complex)
    for a in range(4):
        for b in range(5):
            t[a, b] = x[(5 * a + 4 * b) % 20]
    t = np.fft.fft(t, axis=0)   # DFT4 over a
    t = np.fft.fft(t, axis=1)   # DFT5 over b
    out = np.zeros(20, dtype=complex)
    for ka in range(4):
        for kb in range(5):
            out[(5 * ka + 16 * kb) % 20] = t[ka, kb]
    return out


def slot(r):
    return r if r <= 10 else 30 - r


def model_pair(fa, fb):
    """fa, fb: windowed real frames (400).  Returns power spectra (201,) of each via the kernel's data flow."""
    z = fa + 1j * fb
    Z = np.zeros((21, 20), dtype=complex)            # [slot][n2]
    for c in range(C):                               # step 1 (thread t owns columns 2t, 2t+1)
        y = dft20_pfa(z[C * np.arange(R) + c])       # Y[c][k1]
        y[10] *= np.exp(+2j * np.pi * c / 40.0)      # row 10 is pre-rotated by W_40^-c on the write side
        for k1 in range(R):
            Z[slot(k1), c] = y[k1]
    tw = np.exp(-2j * np.pi * np.outer(np.arange(R + 1), np.arange(C)) / N)   # tw[row][n2]
    PA = np.full(201, np.nan)
    PB = np.full(201, np.nan)
    rows_seen = {}
    for t in range(G):                               # step 3
        a = t
        r1 = a
        r2 = (R - a) if t else R // 2
        s1, s2 = t, 10 + t                           # slots read by thread t
        assert s1 == slot(r1) and s2 == slot(r2)
        X = dft20_pfa(Z[s1] * tw[t])                 # per-lane twiddle registers: W_400^(t*n2)
        D = dft20_pfa(Z[s2] * np.conj(tw[t]))        # second row: conj twiddle, then output rotation by one
        Y = np.roll(D, -1)                           # Y[m] = D[(m+1) % 20]
        lo_base = (R // 2) if t == 0 else a
        hi_base = R if t == 0 else R - a
        for j in range(C):
            if t == 0:
                u = Y[j] if j < C // 2 else X[j]
                v = X[C - j] if j >= C // 2 else Y[C - 1 - j]
            else:
                u, v = X[j], Y[C - 1 - j]
            pa = ((u.real + v.real) ** 2 + (u.imag - v.imag) ** 2) * 0.25
            pb = ((u.imag + v.imag) ** 2 + (u.real - v.real) ** 2) * 0.25
            b = lo_base + R * j if j < C // 2 else hi_base + R * (C - 1 - j)
            row = 10 * j + t
            assert row not in rows_seen
            rows_seen[row] = b
            PA[b], PB[b] = pa, pb
    return PA, PB, rows_seen


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    fa, fb = rng.standard_normal(400), rng.standard_normal(400)
    x = rng.standard_normal(20) + 1j * rng.standard_normal(20)
    print("dft20", np.abs(dft20_pfa(x) - np.fft.fft(x)).max())
    PA, PB, rows = model_pair(fa, fb)
    ra, rb = np.abs(np.fft.fft(fa)[:201]) ** 2, np.abs(np.fft.fft(fb)[:201]) ** 2
    bins = sorted(set(rows.values()))
    print("bins covered", bins[0], bins[-1], len(bins), "missing", sorted(set(range(201)) - set(bins)))
    m = ~np.isnan(PA)
    print("errA", np.abs(PA[m] - ra[m]).max() / ra.max(), "errB", np.abs(PB[m] - rb[m]).max() / rb.max())
