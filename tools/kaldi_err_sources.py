#!/usr/bin/env python
"""Where does the fp32 Kaldi path lose accuracy?  ("compensated arithmetic in the prologue was not tried".)
CPU experiment on JFK with the oracle's semantics (src/fbank.rs:141-236, no CMN so that errors stay per value):
  A  everything f64                                   (the oracle)
  B  prologue (DC removal, pre-emphasis, Povey window) in f32, FFT + projection + ln in f64
  C  prologue in f64, FFT in complex64 (scipy pocketfft), power / projection / ln in f64
  D  prologue f64, FFT f64, power + projection + ln in f32
  E  everything f32 (what the kernel does)
Prints max |x - A| and the share of values beyond 1e-4 / 1e-3 for B..E."""
import os, sys
import numpy as np
import scipy.fft as sfft
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import melspec_oracle as o

x = np.load(os.path.join(ROOT, "tests", "golden", "jfk_pcm_f32.npy"))
L, H, N = 400, 160, 512
T = 1 + (x.size - L) // H
starts = np.arange(T) * H
filt = o.kaldi_mel_filterbank(16000.0, N, 80, 20.0, 8000.0)
win = o.povey_window(L)


def prologue(dt):
    xs = x.astype(dt)
    fr = xs[starts[:, None] + np.arange(L)[None, :]]
    mean = (fr.sum(axis=1, keepdims=True, dtype=dt) / dt(L)).astype(dt)
    z = fr - mean
    y = z.copy()
    y[:, 1:] = z[:, 1:] - dt(0.97) * z[:, :-1]
    prev = np.zeros(T, dt)
    prev[1:] = xs[starts[1:] - 1] - mean[1:, 0]
    y[1:, 0] = z[1:, 0] - dt(0.97) * prev[1:]
    buf = np.zeros((T, N), dt)
    buf[:, :L] = y * win.astype(dt)[None, :]
    return buf


def spectrum(buf, dt):
    if dt == np.float32:
        return sfft.fft(buf.astype(np.complex64), axis=1)[:, :N // 2 + 1]
    return np.fft.fft(buf.astype(np.float64), axis=1)[:, :N // 2 + 1]


def tail(spec, dt):
    re, im = spec.real.astype(dt), spec.imag.astype(dt)
    p = re * re + im * im
    e = (p @ filt.T.astype(dt)).astype(dt)
    return np.log(np.maximum(e, dt(np.finfo(np.float32).eps))).astype(np.float64)


A = tail(spectrum(prologue(np.float64), np.float64), np.float64)
for name, pd, fd, td in (("B f32 prologue only", np.float32, np.float64, np.float64), ("C f32 FFT only", np.float64, np.float32, np.float64),
                         ("D f32 power/projection/ln only", np.float64, np.float64, np.float32), ("E all f32", np.float32, np.float32, np.float32)):
    d = np.abs(tail(spectrum(prologue(pd), fd), td) - A)
    print(f"{name:32s} max {d.max():.2e}  mean {d.mean():.2e}  >1e-4: {100 * (d > 1e-4).mean():.3f} %  >1e-3: {100 * (d > 1e-3).mean():.4f} %")
