"""CPU oracle for the log-mel hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy/f64 restatement of the reference's CPU algorithm (wavey-ai/mel-spec @ ac3bbdd).  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may
import this module; the product path (`mel-spec_b200/`) never does and fails loudly without its
CUDA library.

Parity status: PINNED.  The reference's Rust sources cannot be built here (no cargo/rustc), and its
FFT lives in a third-party crate that is not vendored (`rustfft = "6.2.0"`, Cargo.toml:18, caret
requirement, no lock file); it is used strictly as an unnormalised forward DFT
(src/stft.rs:26-27,79-80,99-110; src/fbank.rs:122-123,193-194), so `numpy.fft.fft` in f64 stands in
for it.  `tests/test_oracle.py` pins this module against every golden vector the reference's own
tests hold for the path (tests/golden/, copied by tests/golden/make_fixtures.py):

  * rust_jfk_golden.npy   (src/rb.rs:134-179, |d| <= 1e-6)        -> whisper_mel_stream, fft 512
  * mel_filters_80x201    (src/mel.rs:837-850, |d| <= 1e-7)       -> slaney_mel_filterbank(16000,400,80)
  * nemo_filters_80x257   (src/mel.rs:852-871, |d| <= 1e-7)       -> slaney_mel_filterbank(16000,512,80)
  * kaldi_fbank_jfk       (src/fbank.rs:439-535, shape + printed distance only in the reference)

Every function cites the reference lines it follows.  Nothing here is copied from the reference; it
is the same arithmetic written as array operations.
"""
from __future__ import annotations

import math

import numpy as np

F32_EPS = float(np.finfo(np.float32).eps)  # f32::EPSILON, src/fbank.rs:213


# --------------------------------------------------------------------------------------------
# Framing / STFT  (src/stft.rs)
# --------------------------------------------------------------------------------------------
def hann_window(fft_size: int) -> np.ndarray:
    """Periodic Hann, f64.  src/stft.rs:141-145 (and 29-31)."""
    i = np.arange(fft_size, dtype=np.float64)
    return 0.5 * (1.0 - np.cos((2.0 * math.pi * i) / float(fft_size)))


def num_frames(n_samples: int, fft_size: int, hop_size: int) -> int:
    """src/stft.rs:153-157: 0 if too short, else (len - fft)/hop + 1 (integer division)."""
    if n_samples < fft_size:
        return 0
    return (n_samples - fft_size) // hop_size + 1


def frame_windows(samples: np.ndarray, fft_size: int, hop_size: int, window: np.ndarray) -> np.ndarray:
    """src/stft.rs:147-169: frame k = samples[k*hop : k*hop+fft] (as f64) * window.  (F, fft) f64."""
    samples = np.asarray(samples, dtype=np.float32)
    nf = num_frames(samples.shape[0], fft_size, hop_size)
    if nf == 0:
        return np.zeros((0, fft_size), dtype=np.float64)
    idx = (np.arange(nf) * hop_size)[:, None] + np.arange(fft_size)[None, :]
    return samples.astype(np.float64)[idx] * window[None, :]


def stft_all(samples: np.ndarray, fft_size: int, hop_size: int) -> np.ndarray:
    """src/stft.rs:89-115 `compute_all_cpu`: full complex forward DFT of each windowed real frame."""
    fr = frame_windows(samples, fft_size, hop_size, hann_window(fft_size))
    if fr.shape[0] == 0:
        return np.zeros((0, fft_size), dtype=np.complex128)
    return np.fft.fft(fr, axis=1)


# --------------------------------------------------------------------------------------------
# Slaney mel filterbank (src/mel.rs:547-643)
# --------------------------------------------------------------------------------------------
def hz_to_mel(f: float, htk: bool = False) -> float:
    """src/mel.rs:591-607."""
    if htk:
        return 2595.0 * math.log10(1.0 + f / 700.0)
    f_sp = 200.0 / 3.0
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    if f >= min_log_hz:
        return min_log_mel + math.log(f / min_log_hz) / logstep
    return f / f_sp


def mel_to_hz(m: float, htk: bool = False) -> float:
    """src/mel.rs:609-625."""
    if htk:
        return 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    f_sp = 200.0 / 3.0
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    if m >= min_log_mel:
        return min_log_hz * math.exp(logstep * (m - min_log_mel))
    return f_sp * m


def mel_frequencies(n_mels: int, fmin: float, fmax: float, htk: bool = False) -> np.ndarray:
    """src/mel.rs:631-637 (ndarray `linspace`: start + i*step, step=(end-start)/(n-1))."""
    lo, hi = hz_to_mel(fmin, htk), hz_to_mel(fmax, htk)
    step = (hi - lo) / (n_mels - 1) if n_mels > 1 else 0.0
    mels = lo + step * np.arange(n_mels, dtype=np.float64)
    return np.array([mel_to_hz(float(m), htk) for m in mels], dtype=np.float64)


def fft_frequencies(sr: float, n_fft: int) -> np.ndarray:
    """src/mel.rs:639-643."""
    return (sr / n_fft) * np.arange(n_fft // 2 + 1, dtype=np.float64)


def slaney_mel_filterbank(sr: float, n_fft: int, n_mels: int, f_min: float | None = None,
                          f_max: float | None = None, htk: bool = False, norm: bool = True) -> np.ndarray:
    """src/mel.rs:547-589 `mel()`: librosa-style triangles + Slaney area normalisation.  (n_mels, n_fft/2+1) f64."""
    fftfreqs = fft_frequencies(sr, n_fft)
    f_min = 0.0 if f_min is None else f_min
    f_max = sr / 2.0 if f_max is None else f_max
    mel_f = mel_frequencies(n_mels + 2, f_min, f_max, htk)
    fdiff = mel_f[1:] - mel_f[:-1]
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, n_fft // 2 + 1), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.minimum(np.clip(lower, 0.0, 1.0), np.clip(upper, 0.0, 1.0))
    if norm:
        enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
        w *= enorm[:, None]
    return w


def sparse_rows(filters: np.ndarray):
    """src/mel.rs:48-71 `SparseMelFilterbank::from_dense`: per row the (bin, weight) pairs with weight != 0."""
    rows = []
    for r in filters:
        nz = np.nonzero(r != 0.0)[0]
        rows.append((nz.astype(np.int64), r[nz].astype(np.float64)))
    return rows


# --------------------------------------------------------------------------------------------
# Whisper path (src/mel.rs:13-32, 148-168, 645-654; src/stft.rs:119-138)
# --------------------------------------------------------------------------------------------
def project_stft_log10(stft: np.ndarray, filters: np.ndarray) -> np.ndarray:
    """src/mel.rs:148-168: E[m] = sum_b W[m,b] |X[b]|^2 with |X|^2 := 0 for b >= len/2; log10(max(E,1e-10)).

    stft: (F, N) complex128; filters: (M, N/2+1) f64  ->  (F, M) f64.
    """
    n = stft.shape[1]
    half = n // 2
    power = np.zeros((stft.shape[0], filters.shape[1]), dtype=np.float64)
    nb = min(half, filters.shape[1])
    power[:, :nb] = stft[:, :nb].real ** 2 + stft[:, :nb].imag ** 2
    # sparse == dense up to summation order (src/mel.rs:887-911 pins them to 1e-12); the sparse rows
    # are contiguous bands summed in ascending bin order, which is what a row-by-row dot does here.
    e = np.zeros((stft.shape[0], filters.shape[0]), dtype=np.float64)
    for m, (bins, wts) in enumerate(sparse_rows(filters)):
        acc = np.zeros(stft.shape[0], dtype=np.float64)
        for b, w in zip(bins, wts):
            acc = acc + w * power[:, b]
        e[:, m] = acc
    return np.log10(np.maximum(e, 1e-10))


def norm_mel_f64(logmel: np.ndarray) -> np.ndarray:
    """src/mel.rs:645-654 (== norm_mel 449-455) applied per frame: t = max - 8; (max(x,t)+4)/4."""
    t = logmel.max(axis=-1, keepdims=True) - 8.0
    return (np.maximum(logmel, t) + 4.0) / 4.0


def norm_mel_vec_f32(logmel_f32: np.ndarray) -> np.ndarray:
    """src/mel.rs:458-469: the f32 variant the reference's GPU backends apply on the host."""
    x = np.asarray(logmel_f32, dtype=np.float32)
    t = x.max(axis=-1, keepdims=True) - np.float32(8.0)
    return ((np.maximum(x, t) + np.float32(4.0)) / np.float32(4.0)).astype(np.float32)


def whisper_mel_batch(samples: np.ndarray, fft_size: int = 400, hop_size: int = 160, n_mels: int = 80,
                      sampling_rate: float = 16000.0, filters: np.ndarray | None = None) -> np.ndarray:
    """src/stft.rs:119-138 `Spectrogram::compute_mel_spectrogram_cpu`.  Returns (F, n_mels) f32, frame-major."""
    if filters is None:
        filters = slaney_mel_filterbank(sampling_rate, fft_size, n_mels)
    st = stft_all(samples, fft_size, hop_size)
    if st.shape[0] == 0:
        return np.zeros((0, n_mels), dtype=np.float32)
    return norm_mel_f64(project_stft_log10(st, filters)).astype(np.float32)


def stream_offset(fft_size: int, hop_size: int) -> int:
    """First-frame offset of the streaming path when fed whole hops (src/stft.rs:61-66): ceil(N/H)*H - N."""
    return -(-fft_size // hop_size) * hop_size - fft_size


def whisper_mel_stream(samples: np.ndarray, fft_size: int = 512, hop_size: int = 160, n_mels: int = 80,
                       sampling_rate: float = 16000.0) -> np.ndarray:
    """RingBuffer::maybe_mel -> Spectrogram::add -> MelSpectrogram::add fed whole hops.

    Literal restatement of src/stft.rs:48-86 (shift hop_buf by one hop, append, emit once idx >= fft) driven as
    src/rb.rs:86-121 drives it (only full hops are ever handed over; a trailing partial hop is never emitted).
    Returns (F, n_mels) f32, frame-major.
    """
    samples = np.asarray(samples, dtype=np.float32)
    window = hann_window(fft_size)
    filters = slaney_mel_filterbank(sampling_rate, fft_size, n_mels)
    hop_buf = np.zeros(fft_size, dtype=np.float64)
    idx = 0
    frames = []
    for h in range(samples.shape[0] // hop_size):
        chunk = samples[h * hop_size:(h + 1) * hop_size].astype(np.float64)
        hop_buf[:fft_size - hop_size] = hop_buf[hop_size:].copy()
        hop_buf[fft_size - hop_size:] = chunk
        idx += hop_size
        if idx >= fft_size:
            frames.append(hop_buf * window)
    if not frames:
        return np.zeros((0, n_mels), dtype=np.float32)
    st = np.fft.fft(np.stack(frames), axis=1)
    return norm_mel_f64(project_stft_log10(st, filters)).astype(np.float32)


def spectrogram_add_mel(chunks, fft_size: int = 400, hop_size: int = 160, n_mels: int = 80, sampling_rate: float = 16000.0):
    """`Spectrogram::add` (src/stft.rs:48-86) called once per element of `chunks`, each followed by `MelSpectrogram::add`
    (src/mel.rs:26-31).  Every chunk has <= hop_size samples (the reference asserts, stft.rs:53); a short chunk is zero-padded
    to a whole hop (stft.rs:56-59); `idx` counts the true samples only (stft.rs:64); a frame is returned once idx >= fft_size
    (stft.rs:66), i.e. with every call from then on.  Returns a list with one (n_mels,) f32 array or None per call."""
    window = hann_window(fft_size)
    filters = slaney_mel_filterbank(sampling_rate, fft_size, n_mels)
    hop_buf = np.zeros(fft_size, dtype=np.float64)
    idx = 0
    out = []
    for c in chunks:
        c = np.asarray(c, dtype=np.float32).astype(np.float64).reshape(-1)
        assert c.size <= hop_size, "frames must be <= hop_size"
        pcm = np.zeros(hop_size, dtype=np.float64)
        pcm[:c.size] = c
        hop_buf[:fft_size - hop_size] = hop_buf[hop_size:].copy()
        hop_buf[fft_size - hop_size:] = pcm
        idx += c.size
        if idx >= fft_size:
            st = np.fft.fft((hop_buf * window)[None, :], axis=1)
            out.append(norm_mel_f64(project_stft_log10(st, filters)).astype(np.float32)[0])
        else:
            out.append(None)
    return out


# --------------------------------------------------------------------------------------------
# Kaldi fbank path (src/fbank.rs)
# --------------------------------------------------------------------------------------------
def kaldi_hz_to_mel(hz: float) -> float:
    """src/fbank.rs:305-307."""
    return 1127.0 * math.log(1.0 + hz / 700.0)


def kaldi_mel_to_hz(mel: float) -> float:
    """src/fbank.rs:311-313."""
    return 700.0 * (math.exp(mel / 1127.0) - 1.0)


def kaldi_mel_filterbank(sample_rate: float = 16000.0, fft_size: int = 512, num_mel_bins: int = 80,
                         low_freq: float = 20.0, high_freq: float = 8000.0) -> np.ndarray:
    """src/fbank.rs:253-301: edges linear in Kaldi-mel, triangles evaluated in Hz, no area norm."""
    nb = fft_size // 2 + 1
    lo, hi = kaldi_hz_to_mel(low_freq), kaldi_hz_to_mel(high_freq)
    hz = [kaldi_mel_to_hz(lo + (hi - lo) * i / (num_mel_bins + 1)) for i in range(num_mel_bins + 2)]
    w = np.zeros((num_mel_bins, nb), dtype=np.float64)
    for m in range(num_mel_bins):
        left, center, right = hz[m], hz[m + 1], hz[m + 2]
        if center <= left or right <= center:
            continue
        for b in range(nb):
            f = b * sample_rate / fft_size
            if left < f <= center:
                w[m, b] = (f - left) / (center - left)
            elif center < f < right:
                w[m, b] = (right - f) / (right - center)
    return w


def povey_window(frame_len: int = 400) -> np.ndarray:
    """src/fbank.rs:100-105."""
    i = np.arange(frame_len, dtype=np.float64)
    a = 2.0 * math.pi * i / float(frame_len - 1)
    return np.power(0.5 - 0.5 * np.cos(a), 0.85)


def kaldi_fbank(samples: np.ndarray, sample_rate: float = 16000.0, num_mel_bins: int = 80,
                frame_length_ms: float = 25.0, frame_shift_ms: float = 10.0, preemphasis: float = 0.97,
                low_freq: float = 20.0, high_freq: float = 0.0, energy_floor: float = 0.0,
                use_log_fbank: bool = True, use_power: bool = True, apply_cmn: bool = True) -> np.ndarray:
    """src/fbank.rs:141-236 `Fbank::compute`.  Returns (T, num_mel_bins) f32."""
    samples = np.asarray(samples, dtype=np.float32)
    frame_len = int(round(frame_length_ms / 1000.0 * sample_rate))      # fbank.rs:68-70
    frame_shift = int(round(frame_shift_ms / 1000.0 * sample_rate))    # fbank.rs:73-75
    fft_size = 1 << (frame_len - 1).bit_length()                       # next_power_of_two, fbank.rs:78-81
    if samples.shape[0] < frame_len:
        return np.zeros((0, num_mel_bins), dtype=np.float32)
    hf = sample_rate / 2.0 if high_freq == 0.0 else high_freq
    filt = kaldi_mel_filterbank(sample_rate, fft_size, num_mel_bins, low_freq, hf)
    window = povey_window(frame_len)
    t = 1 + (samples.shape[0] - frame_len) // frame_shift
    starts = np.arange(t) * frame_shift
    x = samples.astype(np.float64)
    fr = x[starts[:, None] + np.arange(frame_len)[None, :]]            # (T, L)
    mean = fr.sum(axis=1, keepdims=True) / frame_len                   # fbank.rs:166-170 (sequential f64 sum)
    z = fr - mean
    if preemphasis > 0.0:                                              # fbank.rs:172-181
        y = z.copy()
        y[:, 1:] = z[:, 1:] - preemphasis * z[:, :-1]
        prev = np.zeros(t, dtype=np.float64)
        prev[1:] = x[starts[1:] - 1] - mean[1:, 0]
        first = z[:, 0] - preemphasis * prev
        first[0] = z[0, 0] if starts[0] == 0 else first[0]
        y[:, 0] = first
    else:
        y = z
    buf = np.zeros((t, fft_size), dtype=np.float64)
    buf[:, :frame_len] = y * window[None, :]                           # fbank.rs:184-190
    spec = np.fft.fft(buf, axis=1)[:, :fft_size // 2 + 1]
    power = spec.real ** 2 + spec.imag ** 2 if use_power else np.abs(spec)   # fbank.rs:197-203
    e = np.zeros((t, num_mel_bins), dtype=np.float64)
    for m, (bins, wts) in enumerate(sparse_rows(filt)):               # project_power_f64, mel.rs:106-125
        acc = np.zeros(t, dtype=np.float64)
        for b, w in zip(bins, wts):
            acc = acc + w * power[:, b]
        e[:, m] = acc
    floor = energy_floor if energy_floor > 0.0 else F32_EPS            # fbank.rs:210-214
    e = np.maximum(e, floor)
    if use_log_fbank:
        e = np.log(e)
    feats = e.astype(np.float32)                                       # fbank.rs:221
    if apply_cmn and t > 0:                                            # fbank.rs:226-233 (f32 mean per mel)
        feats = feats - feats.mean(axis=0, dtype=np.float32)[None, :]
    return feats.astype(np.float32)


# --------------------------------------------------------------------------------------------
# Synthetic workload: the reference's own 4-tone test signal (src/cuda.rs:494-502) with per-clip
# detune/phase, Gaussian noise, and every 8th clip silent for its first second.
# --------------------------------------------------------------------------------------------
def synth_clip(clip_index: int, n_samples: int, sample_rate: float = 16000.0) -> np.ndarray:
    rng = np.random.default_rng(1234 + clip_index)
    t = np.arange(n_samples, dtype=np.float64) / sample_rate
    x = np.zeros(n_samples, dtype=np.float64)
    for f0, amp in ((220.0, 0.6), (440.0, 0.25), (880.0, 0.10), (1760.0, 0.05)):
        f = f0 * (1.0 + rng.uniform(-0.05, 0.05))
        ph = rng.uniform(0.0, 2.0 * math.pi)
        x += amp * np.sin(2.0 * math.pi * f * t + ph)
    x += 0.01 * rng.standard_normal(n_samples)
    if clip_index % 8 == 7:
        x[:int(sample_rate)] = 0.0
    return x.astype(np.float32)


def reference_test_signal(n_samples: int = 16000, sample_rate: float = 16000.0) -> np.ndarray:
    """The exact f32 signal of src/cuda.rs:494-502 (computed in f32 like the Rust code)."""
    i = np.arange(n_samples, dtype=np.float32)
    t = i / np.float32(sample_rate)
    two_pi = np.float32(2.0) * np.float32(math.pi)
    x = (np.float32(0.6) * np.sin(two_pi * np.float32(220.0) * t)
         + np.float32(0.25) * np.sin(two_pi * np.float32(440.0) * t)
         + np.float32(0.10) * np.sin(two_pi * np.float32(880.0) * t)
         + np.float32(0.05) * np.sin(two_pi * np.float32(1760.0) * t))
    return x.astype(np.float32)


# --------------------------------------------------------------------------------------------
# BatchLogMelSpectrogram — the NeMo/Parakeet-style whole-utterance frontend (src/mel.rs:171-418, 656-756)
# The reference computes this path in f32 (f32 window, rustfft<f32>, f32 projection and ln).  `batch_log_mel` below is
# the same arithmetic in f64 (the semantics); `dtype=np.float32` reruns it with an f32 FFT (scipy pocketfft) to show
# the reference's own f32 noise level.  The reference's tests hold no output vector for this path (only the (128, 101) shape,
# src/mel.rs:943-961, and the filterbank golden nemo_mel_filters.npz, src/mel.rs:852-871, which pins
# slaney_mel_filterbank(16000, 512, 80)).  PINNED instead on the algorithm the reference itself measures this path against
# (README.md:146-158, "Rust-vs-NeMo feature error"): tests/test_nemo_pin.py restates NeMo's FilterbankFeatures independently
# on torch.stft (f64) and holds batch_log_mel to <= 1e-9 of it on the raw and the normalised features, and checks the
# reference's published distance (MAE 0.0012, correlation 0.9997) is met by the f32 rerun of this restatement.
# --------------------------------------------------------------------------------------------
def general_mel_filterbank(sr, n_fft, n_mels, f_min=0.0, f_max=None, htk=False, norm=True):
    """`SparseMelFilterbank::from_mel` -> `mel()` (src/mel.rs:73-85, 547-589) with explicit f_min/f_max/htk/norm."""
    return slaney_mel_filterbank(sr, n_fft, n_mels, f_min, f_max, htk, norm)


def centered_hann_window(n_fft: int, win_length: int, dtype=np.float64) -> np.ndarray:
    """src/mel.rs:708-719: symmetric Hann of win_length centred in n_fft (f32 in the reference)."""
    w = np.zeros(n_fft, dtype=dtype)
    if win_length <= 1:
        return w
    off = (n_fft - win_length) // 2
    i = np.arange(win_length, dtype=dtype)
    w[off:off + win_length] = dtype(0.5) - dtype(0.5) * np.cos((dtype(2.0) * dtype(math.pi) * i) / dtype(win_length - 1.0))
    return w


def batch_num_frames(n: int, n_fft: int, hop: int, center: bool) -> int:
    """src/mel.rs:387-395."""
    if center:
        return n // hop + 1
    return 0 if n < n_fft else (n - n_fft) // hop + 1


def pad_len(n: int, pad_to: int) -> int:
    """src/mel.rs:751-756."""
    return n if pad_to == 0 else -(-n // pad_to) * pad_to


def batch_log_mel(samples, sample_rate=16000, n_fft=512, win_length=400, hop_length=160, n_mels=80, f_min=0.0,
                  f_max=None, htk=False, norm=True, preemphasis=0.0, center=True, log_zero_guard=F32_EPS, pad_to=0,
                  normalize_per_feature=False, dtype=np.float64, out_dtype=np.float32) -> np.ndarray:
    """src/mel.rs:321-385 `compute_flat_with_scratch`.  Returns (n_mels, padded_frames) f32, feature-major
    (`out_dtype=np.float64` keeps the f64 values: used to pin this restatement against an independent torch.stft
    restatement of NeMo's FilterbankFeatures, tests/test_nemo_pin.py)."""
    x = np.asarray(samples, dtype=np.float32)
    if x.size == 0:
        return np.zeros((n_mels, 0), dtype=out_dtype)
    valid = batch_num_frames(x.size, n_fft, hop_length, center)
    padded_frames = pad_len(valid, pad_to)
    wave = x.astype(dtype)
    if preemphasis != 0.0:                                   # src/mel.rs:696-706
        wave = np.concatenate([wave[:1], wave[1:] - dtype(preemphasis) * wave[:-1]])
    if center:                                               # src/mel.rs:685-694
        pad = n_fft // 2
        wave = np.concatenate([np.zeros(pad, dtype), wave, np.zeros(pad, dtype)])
    need = (valid - 1) * hop_length + n_fft if valid else 0
    if wave.size < need:                                     # .get(start+i).unwrap_or(0.0), src/mel.rs:346-347
        wave = np.concatenate([wave, np.zeros(need - wave.size, dtype)])
    window = centered_hann_window(n_fft, win_length, dtype)
    filt = general_mel_filterbank(float(sample_rate), n_fft, n_mels, f_min,
                                  float(sample_rate) / 2.0 if f_max is None else f_max, htk, norm)
    feats = np.zeros((n_mels, padded_frames), dtype=out_dtype)
    if valid:
        idx = (np.arange(valid) * hop_length)[:, None] + np.arange(n_fft)[None, :]
        fr = wave[idx] * window[None, :]
        if dtype == np.float32:
            import scipy.fft as sfft
            spec = sfft.fft(fr.astype(np.complex64), axis=1)[:, :n_fft // 2 + 1]
        else:
            spec = np.fft.fft(fr, axis=1)[:, :n_fft // 2 + 1]
        power = (spec.real ** 2 + spec.imag ** 2).astype(dtype)
        e = np.zeros((valid, n_mels), dtype=dtype)
        for m, (bins, wts) in enumerate(sparse_rows(filt)):            # project_power_f32, src/mel.rs:127-146
            acc = np.zeros(valid, dtype=dtype)
            for b, w in zip(bins, wts):
                acc = acc + dtype(w) * power[:, b]
            e[:, m] = acc
        feats[:, :valid] = np.log(e + dtype(log_zero_guard)).T.astype(out_dtype)
    if normalize_per_feature and valid:                      # src/mel.rs:721-749
        v = feats[:, :valid].astype(dtype)
        mean = v.sum(axis=1, keepdims=True) / dtype(valid)
        denom = max(float(valid) - 1.0, 1.0)
        var = ((v - mean) ** 2).sum(axis=1, keepdims=True) / dtype(denom)
        std = np.sqrt(var) + dtype(1e-5)
        feats[:, :valid] = ((v - mean) / std).astype(out_dtype)
    return feats


# ---------------------------------------------------------------------------------------------------------------
# Output formats: interleave_frames (reference src/mel.rs:480-544) and 8-bit TGA (src/quant.rs:38-165)
# ---------------------------------------------------------------------------------------------------------------
TGA_HEADER_BYTES = 26   # 18-byte header + 8-byte ID field (f32 min, f32 max), src/quant.rs:44-57


def interleave_frames(frames, major_column_order: bool = False, min_width: int = 0) -> np.ndarray:
    """reference src/mel.rs:480-544.  `frames`: (T, n_mels) array ([frame][mel], one column each).  Returns flat f32:
    row-major (n_mels, W) by default, W = T (+1 zero frame if min_width > 0 and T is odd), zero-padded to min_width."""
    fr = np.asarray(frames)
    assert fr.ndim == 2 and fr.shape[0] > 0, "frames is empty"            # src/mel.rs:487
    assert min_width % 2 == 0, "min_width must be even"                   # src/mel.rs:488
    t, m = fr.shape
    if min_width > 0 and t % 2 != 0:                                      # src/mel.rs:497-500
        fr = np.concatenate([fr, np.zeros((1, m), fr.dtype)])
        t += 1
    pad = max(min_width - t, 0)                                           # src/mel.rs:506
    if pad > 0:                                                           # src/mel.rs:509-516
        fr = np.concatenate([fr, np.zeros((pad, m), fr.dtype)])
    if major_column_order:                                                # src/mel.rs:520-530 (frame after frame; the padding
        return fr.astype(np.float32).reshape(-1)                          # block's rows come mel by mel, all zero either way)
    return np.ascontiguousarray(fr.T).astype(np.float32).reshape(-1)      # src/mel.rs:531-541


def _round_half_away(v: np.ndarray) -> np.ndarray:
    """f32::round: half away from zero (numpy's round is half-to-even)."""
    v = np.asarray(v, dtype=np.float32)
    a = np.abs(v)
    r = np.floor(a)
    r = r + ((a - r) >= np.float32(0.5)).astype(np.float32)
    return np.copysign(r, v).astype(np.float32)


def quantize(frame):
    """reference src/quant.rs:140-152, f32 arithmetic step by step.  Returns (u8 array, (min, max))."""
    x = np.asarray(frame, dtype=np.float32).reshape(-1)
    with np.errstate(all="ignore"):
        mn = np.float32(np.fmin.reduce(x, initial=np.float32(np.inf)))
        mx = np.float32(np.fmax.reduce(x, initial=np.float32(-np.inf)))
        scale = np.float32(255.0) / np.float32(mx - mn)
        v = _round_half_away((x - mn).astype(np.float32) * scale)
        v = np.fmin(np.fmax(v, np.float32(0.0)), np.float32(255.0))       # f32::max / f32::min drop NaN
    return v.astype(np.uint8), (float(mn), float(mx))


def dequantize(data, rng) -> np.ndarray:
    """reference src/quant.rs:155-165."""
    mn, mx = np.float32(rng[0]), np.float32(rng[1])
    scale = np.float32(np.float32(mx - mn) / np.float32(255.0))
    return (np.asarray(data, dtype=np.uint8).astype(np.float32) * scale).astype(np.float32) + mn


def tga_8bit_data(data, n_mels: int) -> bytes:
    """reference src/quant.rs:38-64."""
    x = np.asarray(data, dtype=np.float32).reshape(-1)
    q, (mn, mx) = quantize(x)
    width, height = (x.size // n_mels) & 0xFFFF, n_mels & 0xFFFF
    hdr = bytes([8, 0, 3]) + bytes(5) + bytes(4) + int(width).to_bytes(2, "little") + int(height).to_bytes(2, "little") + bytes([8, 0])
    return hdr + np.float32(mn).tobytes() + np.float32(mx).tobytes() + q.tobytes()


def parse_tga_8bit(data: bytes) -> np.ndarray:
    """reference src/quant.rs:66-88."""
    b = bytes(data)
    if len(b) < TGA_HEADER_BYTES:
        raise IOError("failed to fill whole buffer")
    mn, mx = np.frombuffer(b[18:26], dtype="<f4")
    return dequantize(np.frombuffer(b[26:], dtype=np.uint8), (mn, mx))


# ---------------------------------------------------------------------------------------------------------------
# VAD over the mel image: Sobel edge count per column + majority smoothing (reference src/vad.rs:251-486)
# ---------------------------------------------------------------------------------------------------------------
DEFAULT_DETECTION = dict(min_energy=0.98, min_y=11, min_x=5, min_mel=2)      # src/vad.rs:13-22


def vad_raw_classification(img, min_energy: float, min_y: int, min_mel: int) -> np.ndarray:
    """classify_columns_in_frame (src/vad.rs:373-415): img is (height = n_mels, width) row-major; column x is active when
    at least min_y rows y in [min(min_mel, height-2), height-2) have a squared Sobel gradient (src/vad.rs:472-486) of the
    3x3 patch at (y, x) >= min_energy^2.  f64 arithmetic in the reference's operation order.  Returns bool (width-2,)."""
    a = np.asarray(img, dtype=np.float64)
    h, w = a.shape
    if h < 3 or w < 3:                                                     # src/vad.rs:264-266
        return np.zeros(0, dtype=bool)
    if min_y == 0:                                                         # src/vad.rs:382-385
        return np.ones(w - 2, dtype=bool)
    y0 = min(min_mel, h - 2)
    tl, tc, tr = a[y0:h - 2, 0:w - 2], a[y0:h - 2, 1:w - 1], a[y0:h - 2, 2:w]
    ml, mr = a[y0 + 1:h - 1, 0:w - 2], a[y0 + 1:h - 1, 2:w]
    bl, bc, br = a[y0 + 2:h, 0:w - 2], a[y0 + 2:h, 1:w - 1], a[y0 + 2:h, 2:w]
    gx = ((tr + (2.0 * mr)) + br) - ((tl + (2.0 * ml)) + bl)
    gy = ((bl + (2.0 * bc)) + br) - ((tl + (2.0 * tc)) + tr)
    hit = ((gx * gx) + (gy * gy)) >= (min_energy * min_energy)
    return hit.sum(axis=0) >= min_y


def vad_smooth_mask(mask, window: int = 4) -> np.ndarray:
    """smooth_mask (src/vad.rs:343-360): true when at least half of [i-window, i+window] (clipped) is true."""
    m = np.asarray(mask, dtype=bool)
    n = m.size
    pre = np.concatenate([[0], np.cumsum(m)])
    out = np.zeros(n, dtype=bool)
    for i in range(n):
        st, en = max(i - window, 0), min(i + window + 1, n)
        out[i] = (pre[en] - pre[st]) * 2 >= (en - st)
    return out


def vad_boundaries(img, min_energy=0.98, min_y=11, min_x=5, min_mel=2):
    """vad_boundaries (src/vad.rs:251-338) on one (n_mels, width) image.  Returns (non_intersected, intersected) column lists."""
    sm = vad_smooth_mask(vad_raw_classification(img, min_energy, min_y, min_mel), 4)
    idx = np.arange(sm.size)
    return idx[~sm].tolist(), idx[sm].tolist()


def vad_on(intersected, n: int) -> bool:
    """vad_on (src/vad.rs:226-249), including its quirk: a lone first column never counts, so n == 1 needs two columns."""
    if len(intersected) == 0:
        return False
    cnt, prev = 1, intersected[0]
    for ix in intersected[1:]:
        cnt = cnt + 1 if ix == prev + 1 else 1
        if cnt >= n:
            return True
        prev = ix
    return False


def vad_leading_active_columns(intersected) -> int:
    """src/vad.rs:213-224."""
    exp = 0
    for c in intersected:
        if c == exp:
            exp += 1
        elif c > exp:
            break
    return exp


def vad_activity_stream(img, min_energy=0.98, min_y=11, min_x=5, min_mel=2):
    """VoiceActivityDetector::add_activity (src/vad.rs:163-207) fed the columns of img one by one: for every frame index
    i >= min_x - 1 the decision over the window of the last min_x frames.  Returns a list of
    (frame_index, active, leading_active_columns, active_columns, window_columns)."""
    a = np.asarray(img, dtype=np.float64)
    out = []
    for i in range(a.shape[1]):
        if i + 1 < min_x:
            continue
        non, inter = vad_boundaries(a[:, i + 1 - min_x:i + 1], min_energy, min_y, min_x, min_mel)
        out.append((i, bool(inter and inter[0] == 0), vad_leading_active_columns(inter), len(inter), len(inter) + len(non)))
    return out
