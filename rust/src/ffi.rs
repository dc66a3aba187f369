//! `extern "C"` block for include/melspec_b200.h (ABI version 2).  This replaces the reference's private `mod ffi`
//! (src/cuda.rs:161-480: 12 cudart / cuFFT symbols + `launch_mel_kernel`).
//!
//! The `// offset N` comments and the `const _` assertions below state the layout the C header produces on x86-64 / aarch64
//! Linux; tests/test_layout.py compiles a C probe (tests/cpp/layout_probe.c) that prints `offsetof` of every field and compares
//! it with these comments and with the ctypes mirror, so a reordered field fails the build *and* the test-suite.
#![allow(dead_code)]
use std::ffi::{c_char, c_void, CStr};

pub const ABI_VERSION: i32 = 2;
pub const FRONTEND_WHISPER: i32 = 0;
pub const FRONTEND_KALDI: i32 = 1;
pub const FRONTEND_NEMO: i32 = 2;
pub const LAYOUT_FRAME_MAJOR: i32 = 0;
pub const LAYOUT_MEL_MAJOR: i32 = 1;
pub const OK: i32 = 0;
pub const ERR_INVALID_CONFIG: i32 = 1;
pub const ERR_NO_DEVICE: i32 = 2;
pub const ERR_CUDA: i32 = 3;
pub const ERR_INVALID_ARG: i32 = 4;
pub const ERR_UNSUPPORTED: i32 = 5;

#[repr(C)]
pub struct MelspecHandle {
    _private: [u8; 0],
}

#[repr(C)]
pub struct MelspecStream {
    _private: [u8; 0],
}

/// `struct melspec_config` (field order and types must match the header exactly).
#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct MelspecConfig {
    pub frontend: i32,              // offset 0
    pub fft_size: i32,              // offset 4
    pub hop_size: i32,              // offset 8
    pub n_mels: i32,                // offset 12
    pub sampling_rate: f64,         // offset 16
    pub frame_length: i32,          // offset 24
    pub apply_cmn: i32,             // offset 28
    pub use_log_fbank: i32,         // offset 32
    pub use_power: i32,             // offset 36
    pub preemphasis: f64,           // offset 40
    pub low_freq: f64,              // offset 48
    pub high_freq: f64,             // offset 56
    pub energy_floor: f64,          // offset 64
    pub win_length: i32,            // offset 72
    pub center: i32,                // offset 76
    pub pad_to: i32,                // offset 80
    pub normalize_per_feature: i32, // offset 84
    pub htk: i32,                   // offset 88
    pub slaney_norm: i32,           // offset 92
    pub log_zero_guard: f64,        // offset 96
    pub f_min: f64,                 // offset 104
    pub f_max: f64,                 // offset 112
} // sizeof 120

/// `struct melspec_vad_settings` == DetectionSettings (reference src/vad.rs:5-22).
#[repr(C)]
#[derive(Clone, Copy)]
pub struct VadSettings {
    pub min_energy: f64, // offset 0
    pub min_y: i32,      // offset 8
    pub min_x: i32,      // offset 12
    pub min_mel: i32,    // offset 16
} // sizeof 24

const _: () = {
    use std::mem::{offset_of, size_of};
    assert!(size_of::<MelspecConfig>() == 120);
    assert!(offset_of!(MelspecConfig, sampling_rate) == 16);
    assert!(offset_of!(MelspecConfig, frame_length) == 24);
    assert!(offset_of!(MelspecConfig, preemphasis) == 40);
    assert!(offset_of!(MelspecConfig, energy_floor) == 64);
    assert!(offset_of!(MelspecConfig, win_length) == 72);
    assert!(offset_of!(MelspecConfig, slaney_norm) == 92);
    assert!(offset_of!(MelspecConfig, log_zero_guard) == 96);
    assert!(offset_of!(MelspecConfig, f_max) == 112);
    assert!(size_of::<VadSettings>() == 24);
    assert!(offset_of!(VadSettings, min_mel) == 16);
};

#[link(name = "melspec_b200")]
unsafe extern "C" {
    pub fn melspec_abi_version() -> i32;
    pub fn melspec_last_error() -> *const c_char;
    pub fn melspec_default_config(frontend: i32, cfg: *mut MelspecConfig) -> i32;
    pub fn melspec_build_filterbank(cfg: *const MelspecConfig, out: *mut f64, capacity: i64) -> i32;
    pub fn melspec_create(cfg: *const MelspecConfig, device: i32, out: *mut *mut MelspecHandle) -> i32;
    pub fn melspec_destroy(h: *mut MelspecHandle);
    pub fn melspec_num_frames(h: *const MelspecHandle, n_samples: i64) -> i64;
    pub fn melspec_padded_frames(h: *const MelspecHandle, n_samples: i64) -> i64;
    pub fn melspec_max_frames_per_batch(h: *const MelspecHandle) -> i32;
    pub fn melspec_filterbank(h: *const MelspecHandle, out: *mut f64, capacity: i64) -> i32;
    pub fn melspec_compute_device(h: *mut MelspecHandle, d_pcm: *const f32, n_clips: i64, clip_stride: i64, n_samples: i64,
                                  d_lens: *const i32, d_out: *mut f32, out_clip_stride: i64, layout: i32, stream: *mut c_void) -> i32;
    pub fn melspec_compute_host(h: *mut MelspecHandle, h_pcm: *const f32, n_clips: i64, clip_stride: i64, n_samples: i64,
                                h_out: *mut f32, layout: i32, frames_out: *mut i64) -> i32;
    pub fn melspec_compute_host_i16(h: *mut MelspecHandle, h_pcm: *const i16, n_clips: i64, clip_stride: i64, n_samples: i64,
                                    h_out: *mut f32, layout: i32, frames_out: *mut i64) -> i32;
    // streaming: RingBuffer::maybe_mel / Spectrogram::add semantics (reference src/rb.rs:86-121, src/stft.rs:48-86)
    pub fn melspec_stream_create(h: *mut MelspecHandle, max_chunk_samples: i64, out: *mut *mut MelspecStream) -> i32;
    pub fn melspec_stream_push(s: *mut MelspecStream, h_samples: *const f32, n: i64, h_out: *mut f32, out_capacity_frames: i64,
                               frames_emitted: *mut i64) -> i32;
    pub fn melspec_stream_push_hop(s: *mut MelspecStream, h_samples: *const f32, n: i64, h_out_frame: *mut f32, emitted: *mut i32) -> i32;
    pub fn melspec_stream_reset(s: *mut MelspecStream) -> i32;
    pub fn melspec_stream_destroy(s: *mut MelspecStream);
    // output formats: interleave_frames (src/mel.rs:480-544) and the 8-bit TGA quantiser (src/quant.rs:38-165)
    pub fn melspec_interleaved_width(n_frames: i64, min_width: i64) -> i64;
    pub fn melspec_tga_size(n_mels: i32, width: i64) -> i64;
    pub fn melspec_mel_tga_host(h: *mut MelspecHandle, h_pcm: *const f32, n_samples: i64, min_width: i64, h_tga: *mut u8, capacity: i64,
                                width_out: *mut i64, h_img_opt: *mut f32) -> i32;
    pub fn melspec_mel_tga_host_batch(h: *mut MelspecHandle, h_pcm: *const f32, n_clips: i64, clip_stride: i64, n_samples: i64, min_width: i64,
                                      h_tga: *mut u8, tga_stride: i64, width_out: *mut i64) -> i32;
    pub fn melspec_mel_tga_host_batch_i16(h: *mut MelspecHandle, h_pcm: *const i16, n_clips: i64, clip_stride: i64, n_samples: i64,
                                          min_width: i64, h_tga: *mut u8, tga_stride: i64, width_out: *mut i64) -> i32;
    pub fn melspec_quantize_tga_host(h: *mut MelspecHandle, h_img: *const f32, n_mels: i32, width: i64, h_tga: *mut u8) -> i32;
    pub fn melspec_dequantize_tga_host(h: *mut MelspecHandle, h_tga: *const u8, tga_bytes: i64, h_img: *mut f32, capacity: i64) -> i32;
    // VAD over the mel image (src/vad.rs:251-338, 163-207)
    pub fn melspec_vad_default_settings(s: *mut VadSettings) -> i32;
    pub fn melspec_vad_host(h: *mut MelspecHandle, h_img: *const f32, n_mels: i32, width: i64, vs: *const VadSettings,
                            h_smoothed: *mut u8, h_activity_opt: *mut i32) -> i32;
}

/// Copy of the thread-local error text.  The C string is only valid until the next failing call on this thread
/// (include/melspec_b200.h), so it is copied into an owned `String` immediately.
pub fn last_error() -> String {
    unsafe { CStr::from_ptr(melspec_last_error()).to_string_lossy().into_owned() }
}
