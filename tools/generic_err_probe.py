"""Max / mean |gpu - f64 oracle| of the general plan (melspec_generic.cuh) over a spread of fft sizes, including the largest
ones the build accepts, and the loud rejection just above them.  Run on a B200: python tools/generic_err_probe.py"""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import mel_spec_b200 as ms
import melspec_oracle as o
CASES = [(1024, 256, 128, 16000.0), (256, 64, 40, 8000.0), (2048, 512, 80, 44100.0), (4096, 1024, 128, 48000.0),
    (480, 160, 80, 16000.0), (441, 147, 64, 22050.0), (360, 90, 32, 12000.0), (251, 100, 40, 16000.0),
    (400, 320, 80, 16000.0), (512, 128, 80, 16000.0), (512, 160, 80, 22050.0), (16, 4, 4, 16000.0), (8192, 2048, 80, 48000.0),
    (1000, 250, 80, 16000.0), (14, 7, 4, 16000.0), (770, 200, 40, 16000.0), (13652, 4000, 80, 48000.0), (9009, 3000, 64, 48000.0)]
for fft, hop, n_mels, sr in CASES:
    rng = np.random.default_rng(fft * 7 + hop)
    n = max(8 * fft, 20000) + 13
    t = np.arange(n) / sr
    x = (0.5 * np.sin(2 * np.pi * 0.011 * sr * t) + 0.2 * np.sin(2 * np.pi * 0.13 * sr * t) + 0.05 * rng.standard_normal(n)).astype(np.float32)
    x[: n // 5] *= 1e-3
    h = ms.CudaMelSpectrogram(fft, hop, sr, n_mels)
    got = h.compute_mel_spectrogram(x)
    want = o.whisper_mel_batch(x, fft, hop, n_mels, sr)
    d = np.abs(got - want)
    print(fft, hop, n_mels, got.shape, f"max {d.max():.2e} mean {d.mean():.2e}")
    h.close()
try:
    ms.CudaMelSpectrogram(13654, 4000, 48000.0, 80)
    print("13654: created?!")
except ms.CudaError as e:
    print("13654 ->", e.kind, str(e)[:120])
