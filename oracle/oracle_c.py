"""ctypes loader for oracle/libmelspec_oracle.so (the C restatement) — test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _host_tag() -> str:
    """-march=native code must run on the CPU it was built for: tag the build with this host's CPU flags."""
    import hashlib
    try:
        with open("/proc/cpuinfo") as f:
            flags = next((ln for ln in f if ln.startswith("flags")), "")
    except OSError:
        flags = ""
    return hashlib.sha1(flags.encode()).hexdigest()


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libmelspec_oracle.so")
    src = os.path.join(_HERE, "melspec_oracle.c")
    tagf = os.path.join(_HERE, "libmelspec_oracle.so.host")
    tag = _host_tag()
    try:
        same_host = open(tagf).read().strip() == tag
    except OSError:
        same_host = False
    if force or not same_host or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmelspec_oracle.so"], stdout=subprocess.DEVNULL)
        with open(tagf, "w") as f:
            f.write(tag)
    return so


def lib():
    global _LIB
    if _LIB is None:
        try:
            _LIB = C.CDLL(build())
            _probe(_LIB)
        except Exception:  # e.g. built with -march=native on another CPU
            _LIB = C.CDLL(build(force=True))
        L = _LIB
        f32p, f64p = C.POINTER(C.c_float), C.POINTER(C.c_double)
        L.oracle_whisper_batch.restype = C.c_int64
        L.oracle_whisper_batch.argtypes = [f32p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int,
                                           C.c_double, f32p, C.c_int]
        L.oracle_kaldi_batch.restype = C.c_int64
        L.oracle_kaldi_batch.argtypes = [f32p, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_int, C.c_double,
                                         C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, f32p, C.c_int]
        L.oracle_slaney_filterbank.argtypes = [C.c_double, C.c_int, C.c_int, f64p]
        L.oracle_kaldi_filterbank.argtypes = [C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, f64p]
        L.oracle_fft_forward.argtypes = [f64p, f64p, C.c_int]
        L.oracle_num_frames.restype = C.c_int64
        L.oracle_num_frames.argtypes = [C.c_int64, C.c_int, C.c_int]
    return _LIB


def _probe(L):
    L.oracle_num_frames.restype = C.c_int64
    L.oracle_num_frames.argtypes = [C.c_int64, C.c_int, C.c_int]
    assert L.oracle_num_frames(400, 400, 160) == 1


def _f32(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def whisper_batch(pcm: np.ndarray, fft=400, hop=160, n_mels=80, sr=16000.0, threads=1) -> np.ndarray:
    """pcm: (B, S) f32 C-contiguous -> (B, F, n_mels) f32."""
    pcm = np.ascontiguousarray(np.atleast_2d(pcm), dtype=np.float32)
    b, s = pcm.shape
    f = int(lib().oracle_num_frames(s, fft, hop))
    out = np.zeros((b, f, n_mels), dtype=np.float32)
    lib().oracle_whisper_batch(_f32(pcm), b, s, s, fft, hop, n_mels, float(sr), _f32(out), threads)
    return out


def kaldi_batch(pcm: np.ndarray, sr=16000.0, n_mels=80, frame_ms=25.0, shift_ms=10.0, preemph=0.97, low=20.0,
                high=0.0, cmn=True, threads=1) -> np.ndarray:
    pcm = np.ascontiguousarray(np.atleast_2d(pcm), dtype=np.float32)
    b, s = pcm.shape
    fl, sh = int(round(frame_ms / 1000.0 * sr)), int(round(shift_ms / 1000.0 * sr))
    t = 0 if s < fl else 1 + (s - fl) // sh
    out = np.zeros((b, t, n_mels), dtype=np.float32)
    lib().oracle_kaldi_batch(_f32(pcm), b, s, s, float(sr), n_mels, frame_ms, shift_ms, preemph, low, high,
                             1 if cmn else 0, _f32(out), threads)
    return out


def slaney_filterbank(sr, n_fft, n_mels) -> np.ndarray:
    out = np.zeros((n_mels, n_fft // 2 + 1), dtype=np.float64)
    lib().oracle_slaney_filterbank(float(sr), n_fft, n_mels, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def kaldi_filterbank(sr=16000.0, fft=512, n_mels=80, low=20.0, high=8000.0) -> np.ndarray:
    out = np.zeros((n_mels, fft // 2 + 1), dtype=np.float64)
    lib().oracle_kaldi_filterbank(float(sr), fft, n_mels, low, high, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def fft_forward(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.complex128)
    out = np.zeros_like(x)
    lib().oracle_fft_forward(x.view(np.float64).ctypes.data_as(C.POINTER(C.c_double)),
                             out.view(np.float64).ctypes.data_as(C.POINTER(C.c_double)), x.shape[0])
    return out
