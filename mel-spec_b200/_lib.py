"""ctypes binding of the C ABI in include/melspec_b200.h (the same symbols a Rust `extern "C"` block binds).

The shared library is built in-tree by `build()` (nvcc, sm_100a) and is the ONLY compute path: there is no CPU
or PyTorch fallback — if the library is missing or no B200 is present, calls raise `CudaError`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
# MELSPEC_B200_LIB: load another build of the same ABI (A/B measurements of kernel changes; never a fallback)
LIB_PATH = os.environ.get("MELSPEC_B200_LIB") or os.path.join(_PKG, "lib", "libmelspec_b200.so")
SOURCES = [os.path.join(_PKG, "csrc", "melspec_api.cu")]
HEADERS = [os.path.join(_PKG, "csrc", "melspec_kernels.cuh"), os.path.join(_PKG, "csrc", "melspec_generic.cuh"),
           os.path.join(_PKG, "csrc", "melspec_generic2.cuh"),
           os.path.join(_ROOT, "include", "melspec_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-ldl"]

# every symbol include/melspec_b200.h declares
EXPORTS = [
    "melspec_abi_version", "melspec_last_error", "melspec_default_config", "melspec_build_filterbank",
    "melspec_num_frames_cfg", "melspec_create", "melspec_destroy", "melspec_num_frames", "melspec_padded_frames",
    "melspec_max_frames_per_batch", "melspec_n_mels", "melspec_fft_size", "melspec_hop_size", "melspec_filterbank",
    "melspec_compute_device", "melspec_compute_host", "melspec_stream_create", "melspec_stream_push",
    "melspec_stream_reset", "melspec_stream_destroy", "melspec_launch_count",
    "melspec_interleaved_width", "melspec_compute_interleaved_device", "melspec_tga_size", "melspec_quantize_tga_device",
    "melspec_dequantize_tga_device", "melspec_quantize_tga_host", "melspec_dequantize_tga_host", "melspec_mel_tga_host",
    "melspec_vad_default_settings", "melspec_vad_boundaries_device", "melspec_vad_activity_device", "melspec_vad_host",
    "melspec_stream_push_hop", "melspec_compute_host_i16", "melspec_convert_i16_device",
    "melspec_mel_tga_host_batch", "melspec_mel_tga_host_batch_i16",
    "melspec_nccl_unique_id", "melspec_nccl_init", "melspec_gather_nccl", "melspec_nccl_destroy",
]


class MelspecConfig(C.Structure):
    """struct melspec_config (include/melspec_b200.h)."""
    _fields_ = [
        ("frontend", C.c_int32), ("fft_size", C.c_int32), ("hop_size", C.c_int32), ("n_mels", C.c_int32),
        ("sampling_rate", C.c_double),
        ("frame_length", C.c_int32), ("apply_cmn", C.c_int32), ("use_log_fbank", C.c_int32), ("use_power", C.c_int32),
        ("preemphasis", C.c_double), ("low_freq", C.c_double), ("high_freq", C.c_double), ("energy_floor", C.c_double),
        ("win_length", C.c_int32), ("center", C.c_int32), ("pad_to", C.c_int32), ("normalize_per_feature", C.c_int32),
        ("htk", C.c_int32), ("slaney_norm", C.c_int32),
        ("log_zero_guard", C.c_double), ("f_min", C.c_double), ("f_max", C.c_double),
    ]


class VadSettings(C.Structure):
    """struct melspec_vad_settings (include/melspec_b200.h) == DetectionSettings (reference src/vad.rs:5-22)."""
    _fields_ = [("min_energy", C.c_double), ("min_y", C.c_int32), ("min_x", C.c_int32), ("min_mel", C.c_int32)]


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (cross-compiles without a GPU)."""
    if os.environ.get("MELSPEC_B200_LIB"):
        return LIB_PATH
    if force or _stale():
        os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
        nvcc = os.environ.get("NVCC", "nvcc")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
        subprocess.check_call(cmd)
    return LIB_PATH


_LIB = None


def lib() -> C.CDLL:
    """Load the library and declare prototypes.  Raises OSError if it was never built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise OSError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    i32, i64, vp = C.c_int32, C.c_int64, C.c_void_p
    cfgp = C.POINTER(MelspecConfig)
    L.melspec_abi_version.restype = i32
    L.melspec_last_error.restype = C.c_char_p
    L.melspec_default_config.restype = i32
    L.melspec_default_config.argtypes = [i32, cfgp]
    L.melspec_build_filterbank.restype = i32
    L.melspec_build_filterbank.argtypes = [cfgp, C.POINTER(C.c_double), i64]
    L.melspec_num_frames_cfg.restype = i64
    L.melspec_num_frames_cfg.argtypes = [cfgp, i64]
    L.melspec_create.restype = i32
    L.melspec_create.argtypes = [cfgp, i32, C.POINTER(vp)]
    L.melspec_destroy.restype = None
    L.melspec_destroy.argtypes = [vp]
    L.melspec_num_frames.restype = i64
    L.melspec_num_frames.argtypes = [vp, i64]
    L.melspec_padded_frames.restype = i64
    L.melspec_padded_frames.argtypes = [vp, i64]
    for name in ("melspec_max_frames_per_batch", "melspec_n_mels", "melspec_fft_size", "melspec_hop_size"):
        getattr(L, name).restype = i32
        getattr(L, name).argtypes = [vp]
    L.melspec_filterbank.restype = i32
    L.melspec_filterbank.argtypes = [vp, C.POINTER(C.c_double), i64]
    L.melspec_compute_device.restype = i32
    L.melspec_compute_device.argtypes = [vp, vp, i64, i64, i64, vp, vp, i64, i32, vp]
    L.melspec_compute_host.restype = i32
    L.melspec_compute_host.argtypes = [vp, vp, i64, i64, i64, vp, i32, C.POINTER(i64)]
    L.melspec_stream_create.restype = i32
    L.melspec_stream_create.argtypes = [vp, i64, C.POINTER(vp)]
    L.melspec_stream_push.restype = i32
    L.melspec_stream_push.argtypes = [vp, vp, i64, vp, i64, C.POINTER(i64)]
    L.melspec_stream_reset.restype = i32
    L.melspec_stream_reset.argtypes = [vp]
    L.melspec_stream_destroy.restype = None
    L.melspec_stream_destroy.argtypes = [vp]
    L.melspec_launch_count.restype = i64
    L.melspec_launch_count.argtypes = [vp]
    L.melspec_interleaved_width.restype = i64
    L.melspec_interleaved_width.argtypes = [i64, i64]
    L.melspec_compute_interleaved_device.restype = i32
    L.melspec_compute_interleaved_device.argtypes = [vp, vp, i64, i64, i64, i64, vp, i64, vp]
    L.melspec_tga_size.restype = i64
    L.melspec_tga_size.argtypes = [i32, i64]
    L.melspec_quantize_tga_device.restype = i32
    L.melspec_quantize_tga_device.argtypes = [vp, vp, i64, i64, i32, i64, vp, i64, vp]
    L.melspec_dequantize_tga_device.restype = i32
    L.melspec_dequantize_tga_device.argtypes = [vp, vp, i64, i64, i32, i64, vp, i64, vp]
    L.melspec_quantize_tga_host.restype = i32
    L.melspec_quantize_tga_host.argtypes = [vp, vp, i32, i64, vp]
    L.melspec_dequantize_tga_host.restype = i32
    L.melspec_dequantize_tga_host.argtypes = [vp, vp, i64, vp, i64]
    L.melspec_mel_tga_host.restype = i32
    L.melspec_mel_tga_host.argtypes = [vp, vp, i64, i64, vp, i64, C.POINTER(i64), vp]
    for name in ("melspec_mel_tga_host_batch", "melspec_mel_tga_host_batch_i16"):
        getattr(L, name).restype = i32
        getattr(L, name).argtypes = [vp, vp, i64, i64, i64, i64, vp, i64, C.POINTER(i64)]
    vsp = C.POINTER(VadSettings)
    L.melspec_vad_default_settings.restype = i32
    L.melspec_vad_default_settings.argtypes = [vsp]
    L.melspec_vad_boundaries_device.restype = i32
    L.melspec_vad_boundaries_device.argtypes = [vp, vp, i64, i64, i32, i64, vsp, vp, vp, i64, vp]
    L.melspec_vad_activity_device.restype = i32
    L.melspec_vad_activity_device.argtypes = [vp, vp, i64, i64, i32, i64, vsp, vp, i64, vp]
    L.melspec_vad_host.restype = i32
    L.melspec_vad_host.argtypes = [vp, vp, i32, i64, vsp, vp, vp]
    # ABI version 2 (an older build loaded through MELSPEC_B200_LIB for an A/B measurement does not have these)
    if not (os.environ.get("MELSPEC_B200_LIB") and not hasattr(L, "melspec_stream_push_hop")):
        L.melspec_stream_push_hop.restype = i32
        L.melspec_stream_push_hop.argtypes = [vp, vp, i64, vp, C.POINTER(i32)]
        L.melspec_compute_host_i16.restype = i32
        L.melspec_compute_host_i16.argtypes = [vp, vp, i64, i64, i64, vp, i32, C.POINTER(i64)]
        L.melspec_convert_i16_device.restype = i32
        L.melspec_convert_i16_device.argtypes = [vp, vp, i64, i64, i64, vp, i64, vp]
        L.melspec_nccl_unique_id.restype = i32
        L.melspec_nccl_unique_id.argtypes = [vp]
        L.melspec_nccl_init.restype = i32
        L.melspec_nccl_init.argtypes = [vp, vp, i32, i32]
        L.melspec_gather_nccl.restype = i32
        L.melspec_gather_nccl.argtypes = [vp, vp, i64, vp, vp]
        L.melspec_nccl_destroy.restype = i32
        L.melspec_nccl_destroy.argtypes = [vp]
    _LIB = L
    return L


def last_error() -> str:
    return lib().melspec_last_error().decode("utf-8", "replace")
