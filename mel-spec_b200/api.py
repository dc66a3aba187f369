"""Host-side mirror of the reference's public interface for the hot path, over the C ABI.

Same names, argument meaning and error behaviour as the Rust prelude (reference src/prelude.rs:1-23):

    MelConfig                      src/config.rs:1-34
    CudaError                      src/cuda.rs:10-25   (Runtime / Unavailable)
    CudaMelSpectrogram             src/cuda.rs:27-155  new(fft, hop, sr, n_mels), compute_mel_spectrogram, max_frames_per_batch
    Spectrogram.compute_mel_spectrogram   batch semantics of src/stft.rs:119-138 (GPU-backed; there is no CPU path here)
    FbankConfig / Fbank            src/fbank.rs:25-82, 84-250
    mel()                          src/mel.rs:547-589  (filterbank; host-only, no GPU needed)
    RingBuffer                     src/rb.rs:12-122    add_frame / add / maybe_mel, frames come from the streaming C ABI
    interleave_frames              src/mel.rs:480-544  (row-major (n_mels, W) image; produced by the kernel's mel-major store)
    quantize / dequantize / tga_8bit_data / parse_tga_8bit / QuantizationRange      src/quant.rs:5-165
    DetectionSettings / EdgeInfo / vad_boundaries / vad_on / VoiceActivity / VoiceActivityDetector   src/vad.rs:5-338

Arrays are numpy instead of Vec<Vec<f32>> / ndarray; shapes and element order are the reference's.
Everything numeric happens in `lib/libmelspec_b200.so`; nothing here computes features on the CPU.
"""
from __future__ import annotations

import ctypes as C
from collections import deque
from dataclasses import dataclass

import numpy as np

from . import _lib

FRONTEND_WHISPER = 0
FRONTEND_KALDI = 1
FRONTEND_NEMO = 2
LAYOUT_FRAME_MAJOR = 0
LAYOUT_MEL_MAJOR = 1

OK, ERR_INVALID_CONFIG, ERR_NO_DEVICE, ERR_CUDA, ERR_INVALID_ARG, ERR_UNSUPPORTED = range(6)


class CudaError(RuntimeError):
    """reference src/cuda.rs:10-25.  `kind` is 'Unavailable' (construction impossible) or 'Runtime'."""

    def __init__(self, kind: str, msg: str, code: int = ERR_CUDA):
        self.kind, self.msg, self.code = kind, msg, code
        super().__init__(f"CUDA {'unavailable' if kind == 'Unavailable' else 'error'}: {msg}")


def _check(rc: int, constructing: bool = False):
    if rc == OK:
        return
    msg = _lib.last_error()
    if rc in (ERR_NO_DEVICE, ERR_INVALID_CONFIG, ERR_UNSUPPORTED) and constructing:
        raise CudaError("Unavailable", msg, rc)      # src/cuda.rs:45-49, 242-294
    if rc == ERR_INVALID_ARG:
        raise ValueError(msg)
    raise CudaError("Runtime", msg, rc)


@dataclass
class DetectionSettings:
    """reference src/vad.rs:5-22 (Default) and 44-81."""
    min_energy: float = 0.98
    min_y: int = 11
    min_x: int = 5
    min_mel: int = 2

    def _c(self) -> "_lib.VadSettings":
        return _lib.VadSettings(float(self.min_energy), int(self.min_y), int(self.min_x), int(self.min_mel))


@dataclass
class EdgeInfo:
    """reference src/vad.rs:488-520 (gradient_positions is left empty by the reference's current vad_boundaries)."""
    non_intersected_columns: list
    intersected_columns: list

    def non_intersected(self) -> list:
        return list(self.non_intersected_columns)

    def intersected(self) -> list:
        return list(self.intersected_columns)

    def gradient_positions(self) -> set:
        return set()


def vad_on(edge_info: "EdgeInfo", n: int) -> bool:
    """reference src/vad.rs:226-249 (host logic over the device mask), quirk included: the first column alone never fires."""
    cols = edge_info.intersected_columns
    if not cols:
        return False
    cnt, prev = 1, cols[0]
    for ix in cols[1:]:
        cnt = cnt + 1 if ix == prev + 1 else 1
        if cnt >= n:
            return True
        prev = ix
    return False


@dataclass(frozen=True)
class VoiceActivityTimestamps:
    """reference src/vad.rs:117-122."""
    start_ms: int
    center_ms: int
    end_ms: int


@dataclass(frozen=True)
class VadFrameTiming:
    """reference src/vad.rs:90-115."""
    fft_size: int
    hop_size: int
    sampling_rate: float

    def timestamps_for_frame(self, frame_index: int) -> VoiceActivityTimestamps:
        start = frame_index * self.hop_size
        ms = lambda smp: int(np.floor(smp / self.sampling_rate * 1000.0 + 0.5))     # f64::round of a non-negative value
        return VoiceActivityTimestamps(ms(start), ms(start + self.fft_size // 2), ms(start + self.fft_size))


@dataclass(frozen=True)
class VoiceActivity:
    """reference src/vad.rs:124-133."""
    active: bool
    frame_index: int
    leading_active_columns: int
    active_columns: int
    window_columns: int
    confidence: float
    timestamps: "VoiceActivityTimestamps | None" = None


@dataclass
class QuantizationRange:
    """reference src/quant.rs:5-9."""
    min: float
    max: float


@dataclass(frozen=True)
class MelConfig:
    """reference src/config.rs:1-34 (getters become attributes)."""
    fft_size: int
    hop_size: int
    n_mels: int
    sampling_rate: float


@dataclass
class FbankConfig:
    """reference src/fbank.rs:25-64 with its Default."""
    sample_rate: float = 16000.0
    num_mel_bins: int = 80
    frame_length_ms: float = 25.0
    frame_shift_ms: float = 10.0
    dither: float = 0.0
    energy_floor: float = 0.0
    use_energy: bool = False
    use_log_fbank: bool = True
    use_power: bool = True
    preemphasis: float = 0.97
    apply_cmn: bool = True
    low_freq: float = 20.0
    high_freq: float = 0.0

    def frame_length_samples(self) -> int:      # src/fbank.rs:68-70
        return int(round(self.frame_length_ms / 1000.0 * self.sample_rate))

    def frame_shift_samples(self) -> int:       # src/fbank.rs:73-75
        return int(round(self.frame_shift_ms / 1000.0 * self.sample_rate))

    def fft_size(self) -> int:                  # src/fbank.rs:78-81
        n = self.frame_length_samples()
        return 1 << (n - 1).bit_length()


def _whisper_cfg(fft_size, hop_size, n_mels, sampling_rate) -> _lib.MelspecConfig:
    c = _lib.MelspecConfig()
    c.frontend, c.fft_size, c.hop_size, c.n_mels = FRONTEND_WHISPER, int(fft_size), int(hop_size), int(n_mels)
    c.sampling_rate, c.frame_length = float(sampling_rate), int(fft_size)
    return c


def _kaldi_cfg(fc: FbankConfig) -> _lib.MelspecConfig:
    c = _lib.MelspecConfig()
    c.frontend = FRONTEND_KALDI
    c.frame_length, c.hop_size = fc.frame_length_samples(), fc.frame_shift_samples()
    c.fft_size, c.n_mels, c.sampling_rate = fc.fft_size(), int(fc.num_mel_bins), float(fc.sample_rate)
    c.apply_cmn, c.use_log_fbank, c.use_power = int(fc.apply_cmn), int(fc.use_log_fbank), int(fc.use_power)
    c.preemphasis, c.low_freq, c.high_freq, c.energy_floor = fc.preemphasis, fc.low_freq, fc.high_freq, fc.energy_floor
    return c


@dataclass
class BatchLogMelConfig:
    """reference src/mel.rs:171-208 with its Default."""
    sample_rate: int = 16000
    n_fft: int = 512
    win_length: int = 400
    hop_length: int = 160
    n_mels: int = 80
    f_min: float = 0.0
    f_max: float | None = None
    htk: bool = False
    norm: bool = True
    preemphasis: float = 0.0
    center: bool = True
    log_zero_guard: float = float(np.finfo(np.float32).eps)
    pad_to: int = 0
    normalize_per_feature: bool = False


class BatchLogMelError(ValueError):
    """reference src/mel.rs:210-231 (`InvalidConfig`)."""


@dataclass
class BatchLogMelOutput:
    """reference src/mel.rs:233-237: flat feature-major data with its shape."""
    data: np.ndarray
    rows: int
    cols: int


def _nemo_cfg(bc: BatchLogMelConfig) -> _lib.MelspecConfig:
    c = _lib.MelspecConfig()
    c.frontend = FRONTEND_NEMO
    c.fft_size, c.win_length, c.frame_length, c.hop_size = int(bc.n_fft), int(bc.win_length), int(bc.win_length), int(bc.hop_length)
    c.n_mels, c.sampling_rate = int(bc.n_mels), float(bc.sample_rate)
    c.preemphasis, c.center, c.pad_to = float(bc.preemphasis), int(bc.center), int(bc.pad_to)
    c.normalize_per_feature, c.htk, c.slaney_norm = int(bc.normalize_per_feature), int(bc.htk), int(bc.norm)
    c.log_zero_guard, c.f_min, c.f_max = float(bc.log_zero_guard), float(bc.f_min), float(bc.f_max or 0.0)
    return c


def mel(sr: float, n_fft: int, n_mels: int) -> np.ndarray:
    """Slaney filterbank, (n_mels, n_fft/2+1) f64 — `mel(sr, n_fft, n_mels, None, None, false, true)` (src/mel.rs:547-589).
    Built by the C library's host code; needs no GPU."""
    cfg = _whisper_cfg(n_fft, 160, n_mels, sr)
    out = np.zeros((n_mels, n_fft // 2 + 1), dtype=np.float64)
    _check(_lib.lib().melspec_build_filterbank(C.byref(cfg), out.ctypes.data_as(C.POINTER(C.c_double)), out.size), True)
    return out


def kaldi_mel_filterbank(config: FbankConfig | None = None) -> np.ndarray:
    """`Fbank::dense_filterbank` (src/fbank.rs:243-245, 253-301), host-only."""
    fc = config or FbankConfig()
    cfg = _kaldi_cfg(fc)
    out = np.zeros((fc.num_mel_bins, fc.fft_size() // 2 + 1), dtype=np.float64)
    _check(_lib.lib().melspec_build_filterbank(C.byref(cfg), out.ctypes.data_as(C.POINTER(C.c_double)), out.size), True)
    return out


def _ptr(x) -> int:
    """Raw device address of a torch tensor / anything with data_ptr(), or an int."""
    if x is None:
        return 0
    if hasattr(x, "data_ptr"):
        return int(x.data_ptr())
    return int(x)


class _Handle:
    """Owns one melspec_handle (RAII like the reference's Drop, src/cuda.rs:142-148)."""

    def __init__(self, cfg: _lib.MelspecConfig, device: int):
        self._h = C.c_void_p()
        self._L = _lib.lib()
        _check(self._L.melspec_create(C.byref(cfg), int(device), C.byref(self._h)), constructing=True)
        self.n_mels = self._L.melspec_n_mels(self._h)
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.melspec_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- shared entry points -------------------------------------------------------------------
    def num_frames(self, n_samples: int) -> int:
        return int(self._L.melspec_num_frames(self._h, int(n_samples)))

    def launch_count(self) -> int:
        return int(self._L.melspec_launch_count(self._h))

    def compute_device(self, d_pcm, n_clips: int, clip_stride: int, n_samples: int, d_out, *, d_lens=None,
                       out_clip_stride: int = 0, layout: int = LAYOUT_FRAME_MAJOR, stream=None) -> None:
        """`melspec_compute_device`: device-resident PCM in, features out, asynchronous on `stream`.
        `d_pcm`, `d_out`, `d_lens` are torch CUDA tensors (or raw addresses); `stream` a torch.cuda.Stream or handle."""
        s = 0 if stream is None else int(getattr(stream, "cuda_stream", stream))
        _check(self._L.melspec_compute_device(self._h, _ptr(d_pcm), int(n_clips), int(clip_stride), int(n_samples),
                                              _ptr(d_lens), _ptr(d_out), int(out_clip_stride), int(layout), s))

    def compute_host(self, samples, layout: int = LAYOUT_FRAME_MAJOR, out: np.ndarray | None = None) -> np.ndarray:
        """`melspec_compute_host` on a (n_samples,) or (n_clips, n_samples) f32 host array."""
        x = np.asarray(samples, dtype=np.float32)
        single = x.ndim == 1
        x2 = np.ascontiguousarray(x.reshape(1, -1) if single else x)
        n_clips, n = x2.shape
        # the library writes melspec_padded_frames(n) columns per clip (== num_frames(n) except for a NeMo frontend with pad_to,
        # src/mel.rs:751-756), so the buffer is sized with that count
        f = int(self._L.melspec_padded_frames(self._h, int(n)))
        shape = (n_clips, f, self.n_mels) if layout == LAYOUT_FRAME_MAJOR else (n_clips, self.n_mels, f)
        if out is None:
            out = np.zeros(shape, dtype=np.float32)
        if out.dtype != np.float32 or not out.flags.c_contiguous or out.size != int(np.prod(shape)):
            raise ValueError(f"out must be a C-contiguous float32 array of {int(np.prod(shape))} elements (shape {shape})")
        frames = C.c_int64(0)
        _check(self._L.melspec_compute_host(self._h, x2.ctypes.data, n_clips, n, n, out.ctypes.data, int(layout),
                                            C.byref(frames)))
        out = out.reshape(shape)
        return out[0] if single else out

    # ---- output formats (reference src/mel.rs:480-544, src/quant.rs:38-165) ---------------------------------
    def compute_interleaved_device(self, d_pcm, n_clips: int, clip_stride: int, n_samples: int, min_width: int, d_out, *,
                                   out_clip_stride: int = 0, stream=None) -> None:
        """`melspec_compute_interleaved_device`: d_out[clip][mel][W], W = interleaved_width(F, min_width)."""
        s = 0 if stream is None else int(getattr(stream, "cuda_stream", stream))
        _check(self._L.melspec_compute_interleaved_device(self._h, _ptr(d_pcm), int(n_clips), int(clip_stride), int(n_samples),
                                                          int(min_width), _ptr(d_out), int(out_clip_stride), s))

    def quantize_tga_device(self, d_img, n_imgs: int, n_mels: int, width: int, d_tga, *, img_stride: int = 0,
                            tga_stride: int = 0, stream=None) -> None:
        s = 0 if stream is None else int(getattr(stream, "cuda_stream", stream))
        _check(self._L.melspec_quantize_tga_device(self._h, _ptr(d_img), int(n_imgs), int(img_stride), int(n_mels), int(width),
                                                   _ptr(d_tga), int(tga_stride), s))

    def dequantize_tga_device(self, d_tga, n_imgs: int, n_mels: int, width: int, d_img, *, img_stride: int = 0,
                              tga_stride: int = 0, stream=None) -> None:
        s = 0 if stream is None else int(getattr(stream, "cuda_stream", stream))
        _check(self._L.melspec_dequantize_tga_device(self._h, _ptr(d_tga), int(n_imgs), int(tga_stride), int(n_mels), int(width),
                                                     _ptr(d_img), int(img_stride), s))

    def tga_8bit_data(self, data, n_mels: int) -> bytes:
        """src/quant.rs:38-64 on the device: row-major (n_mels, width) f32 image -> TGA bytes."""
        x = np.ascontiguousarray(np.asarray(data, dtype=np.float32).reshape(-1))
        if n_mels <= 0 or x.size == 0 or x.size % n_mels:
            raise ValueError("data length must be a positive multiple of n_mels")
        width = x.size // n_mels
        size = int(self._L.melspec_tga_size(int(n_mels), int(width)))
        if size < 0:      # the reference writes `width as u16` (a wrapped width, src/quant.rs:41); refused here
            raise ValueError("width greater than TARGA max, use [`tga_8bit`]")
        out = np.empty(size, dtype=np.uint8)
        _check(self._L.melspec_quantize_tga_host(self._h, x.ctypes.data, int(n_mels), int(width), out.ctypes.data))
        return out.tobytes()

    def tga_8bit(self, data, n_mels: int) -> list:
        """src/quant.rs:29-36: images wider than a TARGA can hold are cut into strides of u16::MAX columns
        (`chunk_frames_into_strides`, src/quant.rs:100-137: row blocks, then column blocks), one TGA per stride, each with
        its own min / max."""
        x = np.asarray(data, dtype=np.float32).reshape(-1)
        if n_mels <= 0 or x.size % n_mels:
            raise ValueError("data length must be a multiple of n_mels")
        width, stride = x.size // n_mels, 65535
        if width == stride:
            return [self.tga_8bit_data(x, n_mels)]
        img = x.reshape(n_mels, width)
        out = []
        for y in range(0, n_mels, stride):
            for c in range(0, width, stride):
                blk = np.ascontiguousarray(img[y:y + stride, c:c + stride])
                out.append(self.tga_8bit_data(blk, blk.shape[0]))
        return out

    def save_tga_8bit(self, data, n_mels: int, path: str) -> None:
        """src/quant.rs:15-28."""
        x = np.asarray(data, dtype=np.float32).reshape(-1)
        if x.size // n_mels >= 65535:
            raise AssertionError("width greater than TARGA max, use [`tga_8bit`]")      # src/quant.rs:17-21
        with open(path, "wb") as f:
            f.write(self.tga_8bit_data(x, n_mels))

    def load_tga_8bit(self, path: str) -> np.ndarray:
        """src/quant.rs:90-98."""
        with open(path, "rb") as f:
            return self.parse_tga_8bit(f.read())

    def quantize(self, frame):
        """src/quant.rs:140-152: (u8 bytes, QuantizationRange)."""
        x = np.asarray(frame, dtype=np.float32).reshape(-1)
        tga = np.frombuffer(self.tga_8bit_data(x, 1) if x.size <= 65535 else self._quantize_rows(x), dtype=np.uint8)
        rng = QuantizationRange(*np.frombuffer(tga[18:26].tobytes(), dtype="<f4").tolist())
        return tga[26:].copy(), rng

    def _quantize_rows(self, x: np.ndarray) -> bytes:
        # a flat vector longer than a TGA row: use the widest factorisation that fits the u16 fields
        n = x.size
        h = next((k for k in range(2, 65536) if n % k == 0 and n // k <= 65535), None)
        if h is None:
            raise ValueError("vector cannot be laid out as a TGA image")
        return self.tga_8bit_data(x, h)

    def parse_tga_8bit(self, data: bytes) -> np.ndarray:
        """src/quant.rs:66-88 on the device: TGA bytes -> dequantised f32 vector (row-major image)."""
        b = np.frombuffer(bytes(data), dtype=np.uint8)
        if b.size < 26:
            raise IOError("failed to fill whole buffer")
        out = np.empty(b.size - 26, dtype=np.float32)
        _check(self._L.melspec_dequantize_tga_host(self._h, b.ctypes.data, int(b.size), out.ctypes.data, int(out.size)))
        return out

    # ---- VAD over the mel image (reference src/vad.rs:251-486, 163-207) ---------------------------------------
    def vad_boundaries_device(self, d_img, n_imgs: int, n_mels: int, width: int, settings: "DetectionSettings", d_smoothed, *,
                              d_raw=None, img_stride: int = 0, mask_stride: int = 0, stream=None) -> None:
        s = 0 if stream is None else int(getattr(stream, "cuda_stream", stream))
        vs = settings._c()
        _check(self._L.melspec_vad_boundaries_device(self._h, _ptr(d_img), int(n_imgs), int(img_stride), int(n_mels), int(width),
                                                     C.byref(vs), _ptr(d_raw), _ptr(d_smoothed), int(mask_stride), s))

    def vad_activity_device(self, d_raw, n_imgs: int, n_mels: int, width: int, settings: "DetectionSettings", d_activity, *,
                            mask_stride: int = 0, activity_stride: int = 0, stream=None) -> None:
        s = 0 if stream is None else int(getattr(stream, "cuda_stream", stream))
        vs = settings._c()
        _check(self._L.melspec_vad_activity_device(self._h, _ptr(d_raw), int(n_imgs), int(mask_stride), int(n_mels), int(width),
                                                   C.byref(vs), _ptr(d_activity), int(activity_stride), s))

    def vad_boundaries(self, frames, settings: "DetectionSettings | None" = None) -> "EdgeInfo":
        """vad_boundaries (src/vad.rs:251-338) on a (n_mels, width) image (frames side by side, as to_array2 returns)."""
        settings = settings or DetectionSettings()
        a = np.ascontiguousarray(np.asarray(frames, dtype=np.float32))
        if a.ndim != 2:
            raise ValueError("frames must be a (n_mels, width) image")
        h, w = a.shape
        if h < 3 or w < 3:
            return EdgeInfo([], [])
        sm = np.zeros(w - 2, dtype=np.uint8)
        vs = settings._c()
        _check(self._L.melspec_vad_host(self._h, a.ctypes.data, int(h), int(w), C.byref(vs), sm.ctypes.data, 0))
        idx = np.arange(w - 2)
        return EdgeInfo(idx[sm == 0].tolist(), idx[sm != 0].tolist())

    def vad_activities(self, frames, settings: "DetectionSettings | None" = None, timing: "VadFrameTiming | None" = None):
        """VoiceActivityDetector::add_activity (src/vad.rs:163-207) for every column of the image: list of VoiceActivity."""
        settings = settings or DetectionSettings()
        a = np.ascontiguousarray(np.asarray(frames, dtype=np.float32))
        h, w = a.shape
        if w == 0:
            return []
        act = np.zeros((w, 3), dtype=np.int32)
        sm = np.zeros(max(w - 2, 1), dtype=np.uint8)
        vs = settings._c()
        _check(self._L.melspec_vad_host(self._h, a.ctypes.data, int(h), int(w), C.byref(vs), sm.ctypes.data, act.ctypes.data))
        win = max(settings.min_x - 2, 0) if (h >= 3 and settings.min_x >= 3) else 0
        out = []
        for i in range(w):
            if act[i, 0] < 0:
                continue
            out.append(VoiceActivity(bool(act[i, 0]), i, int(act[i, 1]), int(act[i, 2]), win,
                                     (act[i, 2] / win) if win else 0.0, timing.timestamps_for_frame(i) if timing else None))
        return out

    def compute_host_raw(self, h_pcm_ptr: int, n_clips: int, clip_stride: int, n_samples: int, h_out_ptr: int,
                         layout: int = LAYOUT_FRAME_MAJOR) -> int:
        """Pointer form of `melspec_compute_host` (pinned torch host tensors in bench.py)."""
        frames = C.c_int64(0)
        _check(self._L.melspec_compute_host(self._h, int(h_pcm_ptr), int(n_clips), int(clip_stride), int(n_samples),
                                            int(h_out_ptr), int(layout), C.byref(frames)))
        return int(frames.value)


def _i16_methods():
    def compute_host_i16_raw(self, h_pcm_ptr: int, n_clips: int, clip_stride: int, n_samples: int, h_out_ptr: int,
                             layout: int = LAYOUT_FRAME_MAJOR) -> int:
        """Pointer form of `melspec_compute_host_i16` (int16 PCM: half the host-to-device bytes)."""
        frames = C.c_int64(0)
        _check(self._L.melspec_compute_host_i16(self._h, int(h_pcm_ptr), int(n_clips), int(clip_stride), int(n_samples),
                                                int(h_out_ptr), int(layout), C.byref(frames)))
        return int(frames.value)

    def compute_host_i16(self, samples, layout: int = LAYOUT_FRAME_MAJOR) -> np.ndarray:
        """`melspec_compute_host_i16` on a (n_samples,) or (n_clips, n_samples) int16 host array: x / 32768 on the device."""
        x = np.asarray(samples, dtype=np.int16)
        single = x.ndim == 1
        x2 = np.ascontiguousarray(x.reshape(1, -1) if single else x)
        n_clips, n = x2.shape
        f = int(self._L.melspec_padded_frames(self._h, int(n)))
        shape = (n_clips, f, self.n_mels) if layout == LAYOUT_FRAME_MAJOR else (n_clips, self.n_mels, f)
        out = np.zeros(shape, dtype=np.float32)
        self.compute_host_i16_raw(x2.ctypes.data, n_clips, n, n, out.ctypes.data, layout)
        return out[0] if single else out

    _Handle.compute_host_i16_raw = compute_host_i16_raw
    _Handle.compute_host_i16 = compute_host_i16


_i16_methods()


class CudaMelSpectrogram(_Handle):
    """reference src/cuda.rs:27-155.  `CudaMelSpectrogram(fft_size, hop_size, sampling_rate, n_mels)` == `new`."""

    def __init__(self, fft_size: int, hop_size: int, sampling_rate: float, n_mels: int, device: int = 0):
        if fft_size == 0 or hop_size == 0 or n_mels == 0:      # src/cuda.rs:45-49
            raise CudaError("Unavailable", "fft_size, hop_size, and n_mels must be non-zero", ERR_INVALID_CONFIG)
        self.fft_size, self.hop_size, self.sampling_rate = int(fft_size), int(hop_size), float(sampling_rate)
        super().__init__(_whisper_cfg(fft_size, hop_size, n_mels, sampling_rate), device)

    def max_frames_per_batch(self) -> int:                     # src/cuda.rs:84-86, 150-155
        return int(self._L.melspec_max_frames_per_batch(self._h))

    def compute_mel_spectrogram(self, samples) -> np.ndarray:
        """&[f32] -> [frame][mel] f32 (src/cuda.rs:88-101).  Empty / too-short input => shape (0, n_mels)."""
        return self.compute_host(np.asarray(samples, dtype=np.float32).reshape(-1))

    def interleaved_width(self, n_samples: int, min_width: int = 0) -> int:
        return int(self._L.melspec_interleaved_width(self.num_frames(n_samples), int(min_width)))

    def mel_tga(self, samples, min_width: int = 0, return_image: bool = False):
        """PCM -> mel frames -> interleave_frames(.., false, min_width) -> tga_8bit_data, one device pipeline
        (the reference's examples/mel_tga: src/stft.rs + src/mel.rs:480-544 + src/quant.rs:38-64)."""
        x = np.ascontiguousarray(np.asarray(samples, dtype=np.float32).reshape(-1))
        if min_width % 2:
            raise ValueError("min_width must be even")                               # src/mel.rs:488
        f = self.num_frames(x.size)
        if f <= 0:
            raise ValueError("frames is empty")                                      # src/mel.rs:487
        w = int(self._L.melspec_interleaved_width(f, int(min_width)))
        size = int(self._L.melspec_tga_size(self.n_mels, w))
        if size < 0:
            raise ValueError("width greater than TARGA max, use [`tga_8bit`]")
        out = np.empty(size, dtype=np.uint8)
        img = np.empty((self.n_mels, w), dtype=np.float32) if return_image else None
        wout = C.c_int64(0)
        _check(self._L.melspec_mel_tga_host(self._h, x.ctypes.data, int(x.size), int(min_width), out.ctypes.data, size,
                                            C.byref(wout), img.ctypes.data if img is not None else 0))
        return (out.tobytes(), img) if return_image else out.tobytes()

    def mel_tga_batch_raw(self, h_pcm_ptr: int, n_clips: int, clip_stride: int, n_samples: int, h_tga_ptr: int, *, min_width: int = 0,
                          tga_stride: int = 0, int16: bool = False) -> int:
        """Pointer form of `melspec_mel_tga_host_batch[_i16]` (pinned buffers in bench.py); returns the image width."""
        w = C.c_int64(0)
        fn = self._L.melspec_mel_tga_host_batch_i16 if int16 else self._L.melspec_mel_tga_host_batch
        _check(fn(self._h, int(h_pcm_ptr), int(n_clips), int(clip_stride), int(n_samples), int(min_width), int(h_tga_ptr),
                  int(tga_stride), C.byref(w)))
        return int(w.value)

    def mel_tga_batch(self, samples, min_width: int = 0) -> list:
        """A (n_clips, n_samples) batch of f32 or int16 PCM -> one 8-bit TGA image (bytes) per clip, pipelined on the device:
        `tga_8bit_data(interleave_frames(frames, false, min_width))` of every clip (src/mel.rs:480-544, src/quant.rs:38-64)."""
        x = np.asarray(samples)
        i16 = x.dtype == np.int16
        x = np.ascontiguousarray(x if i16 else x.astype(np.float32, copy=False))
        if x.ndim != 2:
            raise ValueError("samples must be (n_clips, n_samples)")
        if min_width % 2:
            raise ValueError("min_width must be even")                               # src/mel.rs:488
        n_clips, n = x.shape
        f = self.num_frames(n)
        if f <= 0:
            raise ValueError("frames is empty")                                      # src/mel.rs:487
        w = int(self._L.melspec_interleaved_width(f, int(min_width)))
        size = int(self._L.melspec_tga_size(self.n_mels, w))
        if size < 0:
            raise ValueError("width greater than TARGA max, use [`tga_8bit`]")
        out = np.empty((n_clips, size), dtype=np.uint8)
        self.mel_tga_batch_raw(x.ctypes.data, n_clips, n, n, out.ctypes.data, min_width=min_width, int16=i16)
        return [out[i].tobytes() for i in range(n_clips)]

    def interleave_frames(self, samples, major_column_order: bool = False, min_width: int = 0) -> np.ndarray:
        """Mel frames of `samples` in the reference's interleave_frames layout (src/mel.rs:480-544), flat f32.
        Row-major (default, what whisper.cpp expects) comes straight from the kernel's mel-major store."""
        if major_column_order:
            # frames one after the other, then the zero frame / padding block (src/mel.rs:520-530): frame-major + zeros
            x = np.asarray(samples, dtype=np.float32).reshape(-1)
            fm = self.compute_host(x)
            if fm.shape[0] == 0:
                raise ValueError("frames is empty")
            if min_width % 2:
                raise ValueError("min_width must be even")
            w = int(self._L.melspec_interleaved_width(fm.shape[0], int(min_width)))
            return np.concatenate([fm.reshape(-1), np.zeros((w - fm.shape[0]) * self.n_mels, np.float32)])
        _, img = self.mel_tga(samples, min_width, return_image=True)
        return img.reshape(-1)


class SpectrogramFrame:
    """What `Spectrogram.add` hands to `MelSpectrogram.add`.  The reference passes the complex FFT frame between the two
    (src/stft.rs:82, src/mel.rs:26); here the whole chain is one fused kernel, so the token already carries the mel frame."""
    __slots__ = ("mel", "fft_size", "n_mels", "sampling_rate")

    def __init__(self, mel, fft_size, n_mels, sampling_rate):
        self.mel, self.fft_size, self.n_mels, self.sampling_rate = mel, fft_size, n_mels, sampling_rate


class Spectrogram:
    """reference src/stft.rs:10-138.  `Spectrogram(fft_size, hop_size).add(frames)` is the streaming overlap-and-save entry
    (src/stft.rs:48-86): <= hop_size samples per call, a short chunk is zero-padded to a whole hop, a frame comes back once
    fft_size true samples have been seen and with every call after that.  `n_mels` / `sampling_rate` default to the Whisper
    values: the device computes STFT and mel in one kernel, so the stream needs them up front.  The classmethod is the batch
    entry (src/stft.rs:119-138), GPU-backed (handles are cached per configuration)."""
    _cache: dict = {}

    def __init__(self, fft_size: int, hop_size: int, n_mels: int = 80, sampling_rate: float = 16000.0, device: int = 0):
        self.fft_size, self.hop_size, self.n_mels, self.sampling_rate = int(fft_size), int(hop_size), int(n_mels), float(sampling_rate)
        self._handle = CudaMelSpectrogram(fft_size, hop_size, sampling_rate, n_mels, device)
        self._L = self._handle._L
        self._s = C.c_void_p()
        _check(self._L.melspec_stream_create(self._handle._h, max(self.hop_size, 1), C.byref(self._s)), constructing=True)
        self._out = np.empty(self.n_mels, dtype=np.float32)

    def add(self, frames):
        x = np.ascontiguousarray(np.asarray(frames, dtype=np.float32).reshape(-1))
        if x.size > self.hop_size:
            raise AssertionError("frames must be <= hop_size")          # src/stft.rs:53
        emitted = C.c_int32(0)
        _check(self._L.melspec_stream_push_hop(self._s, x.ctypes.data if x.size else None, x.size, self._out.ctypes.data,
                                               C.byref(emitted)))
        if not emitted.value:
            return None
        return SpectrogramFrame(self._out.copy(), self.fft_size, self.n_mels, self.sampling_rate)

    def close(self):
        if getattr(self, "_s", None) is not None and self._s:
            self._L.melspec_stream_destroy(self._s)
            self._s = C.c_void_p()
            self._handle.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @classmethod
    def compute_mel_spectrogram(cls, samples, fft_size: int, hop_size: int, n_mels: int, sampling_rate: float,
                                device: int = 0) -> np.ndarray:
        key = (int(fft_size), int(hop_size), int(n_mels), float(sampling_rate), int(device))
        h = cls._cache.get(key)
        if h is None:
            h = cls._cache[key] = CudaMelSpectrogram(fft_size, hop_size, sampling_rate, n_mels, device)
        return h.compute_mel_spectrogram(samples)


class Fbank(_Handle):
    """reference src/fbank.rs:84-250: `Fbank(FbankConfig()).compute(samples) -> (T, num_mel_bins) f32`."""

    def __init__(self, config: FbankConfig | None = None, device: int = 0):
        self.config = config or FbankConfig()
        super().__init__(_kaldi_cfg(self.config), device)

    def compute(self, samples) -> np.ndarray:
        return self.compute_host(np.asarray(samples, dtype=np.float32).reshape(-1))

    def dense_filterbank(self) -> np.ndarray:
        return kaldi_mel_filterbank(self.config)


class BatchLogMelSpectrogram(_Handle):
    """reference src/mel.rs:239-396: `BatchLogMelSpectrogram(BatchLogMelConfig()).compute(samples)` ->
    (n_mels, padded_frames) f32, feature-major; `compute_flat` returns the flat data with rows/cols."""

    def __init__(self, config: BatchLogMelConfig | None = None, device: int = 0):
        self.config = config or BatchLogMelConfig()
        c = self.config
        for bad, msg in ((c.sample_rate == 0, "sample_rate must be > 0"), (c.n_fft == 0, "n_fft must be > 0"),
                         (c.win_length == 0, "win_length must be > 0"), (c.win_length > c.n_fft, "win_length must be <= n_fft"),
                         (c.hop_length == 0, "hop_length must be > 0"), (c.n_mels == 0, "n_mels must be > 0"),
                         (not np.isfinite(c.log_zero_guard) or c.log_zero_guard <= 0.0, "log_zero_guard must be finite and > 0")):
            if bad:
                raise BatchLogMelError(f"invalid log-mel config: {msg}")      # src/mel.rs:656-683
        super().__init__(_nemo_cfg(c), device)

    def padded_frames(self, n_samples: int) -> int:
        return int(self._L.melspec_padded_frames(self._h, int(n_samples)))

    def compute(self, samples) -> np.ndarray:
        x = np.ascontiguousarray(np.asarray(samples, dtype=np.float32).reshape(-1))
        cols = self.padded_frames(x.size)
        out = np.zeros((self.n_mels, cols), dtype=np.float32)
        if cols:
            frames = C.c_int64(0)
            _check(self._L.melspec_compute_host(self._h, x.ctypes.data, 1, x.size, x.size, out.ctypes.data,
                                                LAYOUT_MEL_MAJOR, C.byref(frames)))
        return out

    def compute_flat(self, samples) -> BatchLogMelOutput:
        f = self.compute(samples)
        return BatchLogMelOutput(f.reshape(-1), f.shape[0], f.shape[1])

    def filters(self) -> np.ndarray:
        c = _nemo_cfg(self.config)
        out = np.zeros((self.config.n_mels, self.config.n_fft // 2 + 1), dtype=np.float64)
        _check(_lib.lib().melspec_build_filterbank(C.byref(c), out.ctypes.data_as(C.POINTER(C.c_double)), out.size), True)
        return out


class MelSpectrogram:
    """reference src/mel.rs:13-32: `MelSpectrogram(fft_size, sampling_rate, n_mels).add(fft_frame)` -> (n_mels, 1).  The frame
    token of `Spectrogram.add` already holds the projected, normalised frame (one fused kernel); `add` checks that it was made
    for this configuration and returns it in the reference's shape."""

    def __init__(self, fft_size: int, sampling_rate: float, n_mels: int):
        self.fft_size, self.sampling_rate, self.n_mels = int(fft_size), float(sampling_rate), int(n_mels)

    def add(self, fft: "SpectrogramFrame") -> np.ndarray:
        if (fft.fft_size, fft.n_mels, fft.sampling_rate) != (self.fft_size, self.n_mels, self.sampling_rate):
            raise ValueError("frame was produced by a Spectrogram with a different fft_size / n_mels / sampling_rate")
        return fft.mel.reshape(-1, 1)


class RingBuffer:
    """reference src/rb.rs:12-122 on the device: samples are queued on the host (bounded FIFO that drops the oldest
    samples when full, like the VecDeque build of the reference), whole hops are pushed to the streaming C ABI, and
    `maybe_mel()` hands out one (n_mels, 1) frame at a time (f32; the reference's Array2<f64> cast to f32 at the end)."""

    def __init__(self, config: MelConfig, capacity: int, device: int = 0, max_chunk_samples: int = 1 << 16):
        self.config = config
        self.capacity = int(capacity)
        self._fifo: deque = deque()
        self._fifo_len = 0
        self._handle = CudaMelSpectrogram(config.fft_size, config.hop_size, config.sampling_rate, config.n_mels, device)
        self._L = self._handle._L
        self._s = C.c_void_p()
        self._max_chunk = int(max_chunk_samples) // config.hop_size * config.hop_size
        _check(self._L.melspec_stream_create(self._handle._h, self._max_chunk, C.byref(self._s)), constructing=True)
        self._ready: deque = deque()
        self._out = np.empty((self._max_chunk // config.hop_size + 4, config.n_mels), dtype=np.float32)

    def close(self):
        if getattr(self, "_s", None) is not None and self._s:
            self._L.melspec_stream_destroy(self._s)
            self._s = C.c_void_p()
        self._handle.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_frame(self, samples) -> None:                      # src/rb.rs:54-70
        x = np.asarray(samples, dtype=np.float32).reshape(-1)
        if x.size > self.capacity:
            x = x[-self.capacity:]
        over = self._fifo_len + x.size - self.capacity
        while over > 0 and self._fifo:
            head = self._fifo[0]
            if head.size <= over:
                self._fifo.popleft(); self._fifo_len -= head.size; over -= head.size
            else:
                self._fifo[0] = head[over:]; self._fifo_len -= over; over = 0
        self._fifo.append(x.copy())
        self._fifo_len += x.size

    def add(self, sample: float) -> None:                      # src/rb.rs:72-84
        self.add_frame(np.asarray([sample], dtype=np.float32))

    def _drain_hops(self, max_hops: int | None = None) -> None:
        hop = self.config.hop_size
        n_hops = self._fifo_len // hop
        if max_hops is not None:
            n_hops = min(n_hops, max_hops)
        n_hops = min(n_hops, self._max_chunk // hop)
        if n_hops == 0:
            return
        need = n_hops * hop
        parts, got = [], 0
        while got < need:
            head = self._fifo.popleft()
            take = min(head.size, need - got)
            parts.append(head[:take])
            if take < head.size:
                self._fifo.appendleft(head[take:])
            got += take
        self._fifo_len -= need
        chunk = np.ascontiguousarray(np.concatenate(parts))
        emitted = C.c_int64(0)
        _check(self._L.melspec_stream_push(self._s, chunk.ctypes.data, chunk.size, self._out.ctypes.data,
                                           self._out.shape[0], C.byref(emitted)))
        for k in range(int(emitted.value)):
            self._ready.append(self._out[k].copy())

    def maybe_mel(self):                                       # src/rb.rs:86-121
        """One hop of queued samples -> at most one frame, exactly like the reference: returns an (n_mels, 1) array
        or None (not enough samples for a hop yet, or the stream has not seen fft_size samples)."""
        if not self._ready:
            if self._fifo_len < self.config.hop_size:
                return None
            self._drain_hops(max_hops=1)
        if not self._ready:
            return None
        return self._ready.popleft().reshape(-1, 1)

    def drain(self) -> np.ndarray:
        """Convenience for long streams: push every queued whole hop in large chunks, return all frames (F, n_mels)."""
        while self._fifo_len >= self.config.hop_size:
            self._drain_hops()
        out = np.stack(list(self._ready)) if self._ready else np.zeros((0, self.config.n_mels), np.float32)
        self._ready.clear()
        return out
