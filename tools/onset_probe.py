"""Worst-case numerics probe: a quiet (but non-zero) stretch next to a loud one.  The specialised kernels transform two
neighbouring frames as the real / imaginary part of one complex FFT, so the fp32 rounding noise of the louder frame leaks
into the quieter one; the general plan transforms every frame on its own.  Prints max|gpu - oracle| per level ratio."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import mel_spec_b200 as ms
import melspec_oracle as o

sr = 16000.0
n = 40000
rng = np.random.default_rng(0)
t = np.arange(n) / sr
base = (0.5 * np.sin(2 * np.pi * 176 * t) + 0.2 * np.sin(2 * np.pi * 2080 * t) + 0.05 * rng.standard_normal(n)).astype(np.float32)
for fft, hop in ((400, 160), (512, 160), (400, 320), (1024, 256)):
    h = ms.CudaMelSpectrogram(fft, hop, sr, 80)
    for q in (1.0, 1e-1, 1e-2, 1e-3, 1e-4, 1e-5):
        x = base.copy()
        x[: n // 5] *= q
        x[3 * n // 5: 4 * n // 5] *= q
        got = h.compute_mel_spectrogram(x)
        want = o.whisper_mel_batch(x, fft, hop, 80, sr)
        d = np.abs(got - want)
        print(f"fft {fft} hop {hop} quiet/loud {q:7.0e}: max {d.max():.2e}  mean {d.mean():.2e}  frames>1e-4: {(d.max(axis=1) > 1e-4).sum()} of {d.shape[0]}")
    h.close()
