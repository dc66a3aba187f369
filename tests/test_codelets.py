"""CPU check of the DFT codelets the CUDA kernels are built from: the scalar 20/16/32-point codelets of
mel-spec_b200/csrc/melspec_kernels.cuh are extracted and compiled for the host with g++ (the __device__ qualifiers
are defined away) and compared with a direct O(N^2) DFT in f64.  The packed FADD2/FFMA2 variants only exist on the
device and are covered by the GPU parity tests."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r'''
int main() {
    double worst = 0;
    srand(7);
    for (int trial = 0; trial < 8; ++trial) {
        { float xr[32], xi[32]; std::complex<double> x[32];
          for (int n = 0; n < 32; ++n) { xr[n] = (float)rand() / RAND_MAX - 0.5f; xi[n] = (float)rand() / RAND_MAX - 0.5f; x[n] = {xr[n], xi[n]}; }
          dft32(xr, xi);
          for (int k = 0; k < 32; ++k) { std::complex<double> a = 0; for (int n = 0; n < 32; ++n) a += x[n] * std::polar(1.0, -2 * M_PI * n * k / 32.0);
            worst = std::max(worst, std::abs(a - std::complex<double>(xr[k], xi[k]))); } }
        { float xr[16], xi[16]; std::complex<double> x[16];
          for (int n = 0; n < 16; ++n) { xr[n] = (float)rand() / RAND_MAX - 0.5f; xi[n] = (float)rand() / RAND_MAX - 0.5f; x[n] = {xr[n], xi[n]}; }
          dft16<1, 16>(xr, xi, 0);
          for (int k = 0; k < 16; ++k) { std::complex<double> a = 0; for (int n = 0; n < 16; ++n) a += x[n] * std::polar(1.0, -2 * M_PI * n * k / 16.0);
            worst = std::max(worst, std::abs(a - std::complex<double>(xr[k], xi[k]))); } }
        { float xr[20], xi[20]; std::complex<double> x[20];
          for (int n = 0; n < 20; ++n) { xr[n] = (float)rand() / RAND_MAX - 0.5f; xi[n] = (float)rand() / RAND_MAX - 0.5f; x[n] = {xr[n], xi[n]}; }
          dft20(xr, xi);
          for (int k = 0; k < 20; ++k) { std::complex<double> a = 0; for (int n = 0; n < 20; ++n) a += x[n] * std::polar(1.0, -2 * M_PI * n * k / 20.0);
            worst = std::max(worst, std::abs(a - std::complex<double>(xr[k], xi[k]))); } }
    }
    printf("%g\n", worst);
    return worst < 2e-6 ? 0 : 1;
}
'''


def test_scalar_codelets_match_direct_dft(tmp_path):
    src = open(os.path.join(ROOT, "mel-spec_b200", "csrc", "melspec_kernels.cuh")).read()
    d5 = src[src.index("#define MS_C1"):src.index("// ---- packed (two transforms at once) codelets")]
    p2 = src[src.index("// ---- power-of-two codelets"):src.index("// ------------------------------------------------------------------------------------------------ plan-400 constants")]
    cpp = ("#include <cstdio>\n#include <cmath>\n#include <complex>\n#include <cstdlib>\n#include <algorithm>\n"
           "#define __device__\n#define __forceinline__ inline\n" + d5 + p2 + HARNESS)
    f = tmp_path / "codelets.cpp"
    f.write_text(cpp)
    exe = tmp_path / "codelets"
    subprocess.check_call(["g++", "-O2", "-o", str(exe), str(f)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout


def test_plan_index_algebra():
    """tools/model_plan.py: the Cooley-Tukey / conjugate-twiddle / untangle / power-row algebra both kernels hard-code."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import numpy as np
    import model_plan as mp
    for R, C in ((20, 20), (32, 16)):
        n = R * C
        rng = np.random.default_rng(R)
        fa, fb = rng.standard_normal(n), rng.standard_normal(n)
        pa, pb, rows = mp.model_pair(fa, fb, R, C)
        ra, rb = np.abs(np.fft.fft(fa)[:n // 2 + 1]) ** 2, np.abs(np.fft.fft(fb)[:n // 2 + 1]) ** 2
        assert sorted(set(rows.values())) == list(range(1, n // 2 + 1))          # every bin but DC, exactly once
        assert np.abs(pa[1:] - ra[1:]).max() / ra.max() < 1e-12 and np.abs(pb[1:] - rb[1:]).max() / rb.max() < 1e-12
        assert all(mp.row_of_bin(b, R, C) == r for r, b in rows.items())
