"""melspec_b200 — B200-native log-mel frontend behind the wavey-ai/mel-spec prelude names.

Layout: csrc/ (the sm_100a kernel + the C ABI), lib/ (the built shared library, git-ignored), _lib.py (ctypes
binding of include/melspec_b200.h), api.py (host-side mirror of the reference interface).
"""
from ._lib import build, lib, last_error, LIB_PATH, EXPORTS  # noqa: F401
from .api import (  # noqa: F401
    BatchLogMelConfig, BatchLogMelError, BatchLogMelOutput, BatchLogMelSpectrogram,
    CudaError, CudaMelSpectrogram, DetectionSettings, EdgeInfo, Fbank, FbankConfig, MelConfig, QuantizationRange, RingBuffer,
    Spectrogram, SpectrogramFrame, MelSpectrogram, VadFrameTiming, VoiceActivity, VoiceActivityTimestamps, vad_on,
    kaldi_mel_filterbank, mel,
    FRONTEND_KALDI, FRONTEND_NEMO, FRONTEND_WHISPER, LAYOUT_FRAME_MAJOR, LAYOUT_MEL_MAJOR,
)

__all__ = [
    "BatchLogMelConfig", "BatchLogMelError", "BatchLogMelOutput", "BatchLogMelSpectrogram",
    "CudaError", "CudaMelSpectrogram", "DetectionSettings", "EdgeInfo", "VadFrameTiming", "VoiceActivity", "vad_on", "Fbank", "FbankConfig", "MelConfig", "QuantizationRange", "RingBuffer", "Spectrogram", "SpectrogramFrame", "MelSpectrogram",
    "kaldi_mel_filterbank", "mel", "build", "lib",
]
