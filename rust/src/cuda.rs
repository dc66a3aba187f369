//! Drop-in replacement for the reference's `src/cuda.rs` (wavey-ai/mel-spec): same public items
//! (`CudaError`, `CudaMelSpectrogram::{new, compute_mel_spectrogram, max_frames_per_batch}`), but the private
//! `mod ffi` now binds the melspec_b200 C ABI (`include/melspec_b200.h`) instead of cudart + cuFFT +
//! `launch_mel_kernel` (reference src/cuda.rs:185-220).
//!
//! SOURCE ONLY: this image has no cargo/rustc, so this file is not compiled or tested here; everything it calls
//! is exercised through the same C ABI from Python (`tests/`) — see INTEGRATION.md.
use std::ffi::{c_char, c_void, CStr};
use std::ptr;

#[derive(Debug)]
pub enum CudaError {
    Runtime(String),
    Unavailable(String),
}

impl std::fmt::Display for CudaError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        match self {
            Self::Runtime(msg) => write!(f, "CUDA error: {msg}"),
            Self::Unavailable(msg) => write!(f, "CUDA unavailable: {msg}"),
        }
    }
}

impl std::error::Error for CudaError {}

pub struct CudaMelSpectrogram {
    handle: *mut ffi::MelspecHandle, // raw pointer => !Send + !Sync, like the reference struct
    n_mels: usize,
}

impl CudaMelSpectrogram {
    pub fn new(fft_size: usize, hop_size: usize, sampling_rate: f64, n_mels: usize) -> Result<Self, CudaError> {
        if fft_size == 0 || hop_size == 0 || n_mels == 0 {
            return Err(CudaError::Unavailable("fft_size, hop_size, and n_mels must be non-zero".into()));
        }
        let mut cfg = ffi::MelspecConfig::default();
        unsafe { ffi::melspec_default_config(ffi::FRONTEND_WHISPER, &mut cfg) };
        cfg.fft_size = fft_size as i32;
        cfg.hop_size = hop_size as i32;
        cfg.n_mels = n_mels as i32;
        cfg.sampling_rate = sampling_rate;
        let mut handle = ptr::null_mut();
        let rc = unsafe { ffi::melspec_create(&cfg, 0, &mut handle) };
        if rc != 0 {
            return Err(CudaError::Unavailable(ffi::last_error()));
        }
        Ok(Self { handle, n_mels })
    }

    pub fn max_frames_per_batch(&self) -> usize {
        unsafe { ffi::melspec_max_frames_per_batch(self.handle) as usize }
    }

    /// `&[f32]` -> `[frame][mel]`, the reference's signature (src/cuda.rs:88-101).
    pub fn compute_mel_spectrogram(&mut self, samples: &[f32]) -> Result<Vec<Vec<f32>>, CudaError> {
        let frames = unsafe { ffi::melspec_num_frames(self.handle, samples.len() as i64) } as usize;
        if frames == 0 {
            return Ok(Vec::new());
        }
        let mut flat = vec![0.0f32; frames * self.n_mels];
        let rc = unsafe {
            ffi::melspec_compute_host(
                self.handle,
                samples.as_ptr(),
                1,
                samples.len() as i64,
                samples.len() as i64,
                flat.as_mut_ptr(),
                ffi::LAYOUT_FRAME_MAJOR,
                ptr::null_mut(),
            )
        };
        if rc != 0 {
            return Err(CudaError::Runtime(ffi::last_error()));
        }
        Ok(flat.chunks(self.n_mels).map(|row| row.to_vec()).collect())
    }

    /// Device-resident batch entry (no host copies): `d_pcm` / `d_out` are CUDA device pointers.
    ///
    /// # Safety
    /// The pointers must be valid device allocations of the sizes described in `include/melspec_b200.h`.
    pub unsafe fn compute_device(
        &mut self,
        d_pcm: *const f32,
        n_clips: usize,
        clip_stride: usize,
        n_samples: usize,
        d_out: *mut f32,
        stream: *mut c_void,
    ) -> Result<(), CudaError> {
        let rc = ffi::melspec_compute_device(
            self.handle,
            d_pcm,
            n_clips as i64,
            clip_stride as i64,
            n_samples as i64,
            ptr::null(),
            d_out,
            0,
            ffi::LAYOUT_FRAME_MAJOR,
            stream,
        );
        if rc != 0 {
            return Err(CudaError::Runtime(ffi::last_error()));
        }
        Ok(())
    }
}

impl CudaMelSpectrogram {
    /// `interleave_frames(frames, false, min_width)` (reference src/mel.rs:480-544) + `tga_8bit_data` (src/quant.rs:38-64)
    /// of the mel frames of `samples`, computed on the device in one pipeline.  Returns (tga bytes, width).
    pub fn mel_tga(&mut self, samples: &[f32], min_width: usize) -> Result<(Vec<u8>, usize), CudaError> {
        assert!(min_width % 2 == 0, "min_width must be even");
        let frames = unsafe { ffi::melspec_num_frames(self.handle, samples.len() as i64) };
        assert!(frames > 0, "frames is empty");
        let width = unsafe { ffi::melspec_interleaved_width(frames, min_width as i64) };
        let size = unsafe { ffi::melspec_tga_size(self.n_mels as i32, width) };
        assert!(size > 0, "width greater than TARGA max, use [`tga_8bit`]");
        let mut out = vec![0u8; size as usize];
        let mut w = 0i64;
        let rc = unsafe {
            ffi::melspec_mel_tga_host(self.handle, samples.as_ptr(), samples.len() as i64, min_width as i64, out.as_mut_ptr(),
                                      size, &mut w, ptr::null_mut())
        };
        if rc != 0 {
            return Err(CudaError::Runtime(ffi::last_error()));
        }
        Ok((out, w as usize))
    }

    /// `vad_boundaries` (reference src/vad.rs:251-338) on a row-major (n_mels, width) image: the smoothed per-column mask
    /// (`true` = column in `EdgeInfo::intersected()`), computed by the device kernel in f64 like the reference.
    pub fn vad_mask(&mut self, image: &[f32], settings: (f64, usize, usize, usize)) -> Result<Vec<bool>, CudaError> {
        let width = image.len() / self.n_mels;
        if self.n_mels < 3 || width < 3 {
            return Ok(Vec::new());
        }
        let vs = ffi::VadSettings { min_energy: settings.0, min_y: settings.1 as i32, min_x: settings.2 as i32, min_mel: settings.3 as i32 };
        let mut mask = vec![0u8; width - 2];
        let rc = unsafe {
            ffi::melspec_vad_host(self.handle, image.as_ptr(), self.n_mels as i32, width as i64, &vs, mask.as_mut_ptr(), ptr::null_mut())
        };
        if rc != 0 {
            return Err(CudaError::Runtime(ffi::last_error()));
        }
        Ok(mask.into_iter().map(|b| b != 0).collect())
    }
}

impl Drop for CudaMelSpectrogram {
    fn drop(&mut self) {
        unsafe { ffi::melspec_destroy(self.handle) };
    }
}

mod ffi {
    use super::{c_char, c_void, CStr};

    pub const FRONTEND_WHISPER: i32 = 0;
    pub const LAYOUT_FRAME_MAJOR: i32 = 0;

    #[repr(C)]
    pub struct MelspecHandle {
        _private: [u8; 0],
    }

    /// `struct melspec_config` of include/melspec_b200.h (field order and types must match exactly).
    #[repr(C)]
    #[derive(Default, Clone, Copy)]
    pub struct MelspecConfig {
        pub frontend: i32,
        pub fft_size: i32,
        pub hop_size: i32,
        pub n_mels: i32,
        pub sampling_rate: f64,
        pub frame_length: i32,
        pub apply_cmn: i32,
        pub use_log_fbank: i32,
        pub use_power: i32,
        pub preemphasis: f64,
        pub low_freq: f64,
        pub high_freq: f64,
        pub energy_floor: f64,
        // NeMo block (BatchLogMelConfig, reference src/mel.rs:171-208)
        pub win_length: i32,
        pub center: i32,
        pub pad_to: i32,
        pub normalize_per_feature: i32,
        pub htk: i32,
        pub slaney_norm: i32,
        pub log_zero_guard: f64,
        pub f_min: f64,
        pub f_max: f64,
    }

    #[repr(C)]
    pub struct MelspecStream {
        _private: [u8; 0],
    }

    /// `struct melspec_vad_settings` == DetectionSettings (reference src/vad.rs:5-22).
    #[repr(C)]
    #[derive(Clone, Copy)]
    pub struct VadSettings {
        pub min_energy: f64,
        pub min_y: i32,
        pub min_x: i32,
        pub min_mel: i32,
    }

    #[link(name = "melspec_b200")]
    unsafe extern "C" {
        pub fn melspec_default_config(frontend: i32, cfg: *mut MelspecConfig) -> i32;
        pub fn melspec_create(cfg: *const MelspecConfig, device: i32, out: *mut *mut MelspecHandle) -> i32;
        pub fn melspec_destroy(h: *mut MelspecHandle);
        pub fn melspec_num_frames(h: *const MelspecHandle, n_samples: i64) -> i64;
        pub fn melspec_max_frames_per_batch(h: *const MelspecHandle) -> i32;
        pub fn melspec_compute_device(
            h: *mut MelspecHandle,
            d_pcm: *const f32,
            n_clips: i64,
            clip_stride: i64,
            n_samples: i64,
            d_lens: *const i32,
            d_out: *mut f32,
            out_clip_stride: i64,
            layout: i32,
            stream: *mut c_void,
        ) -> i32;
        pub fn melspec_compute_host(
            h: *mut MelspecHandle,
            h_pcm: *const f32,
            n_clips: i64,
            clip_stride: i64,
            n_samples: i64,
            h_out: *mut f32,
            layout: i32,
            frames_out: *mut i64,
        ) -> i32;
        pub fn melspec_last_error() -> *const c_char;
        // streaming: RingBuffer::maybe_mel / Spectrogram::add semantics (reference src/rb.rs:86-121, src/stft.rs:48-86)
        pub fn melspec_stream_create(h: *mut MelspecHandle, max_chunk_samples: i64, out: *mut *mut MelspecStream) -> i32;
        pub fn melspec_stream_push(
            s: *mut MelspecStream,
            h_samples: *const f32,
            n: i64,
            h_out: *mut f32,
            out_capacity_frames: i64,
            frames_emitted: *mut i64,
        ) -> i32;
        pub fn melspec_stream_reset(s: *mut MelspecStream) -> i32;
        pub fn melspec_stream_destroy(s: *mut MelspecStream);
        // output formats: interleave_frames (src/mel.rs:480-544) and the 8-bit TGA quantiser (src/quant.rs:38-165)
        pub fn melspec_interleaved_width(n_frames: i64, min_width: i64) -> i64;
        pub fn melspec_tga_size(n_mels: i32, width: i64) -> i64;
        pub fn melspec_mel_tga_host(
            h: *mut MelspecHandle,
            h_pcm: *const f32,
            n_samples: i64,
            min_width: i64,
            h_tga: *mut u8,
            capacity: i64,
            width_out: *mut i64,
            h_img_opt: *mut f32,
        ) -> i32;
        pub fn melspec_quantize_tga_host(h: *mut MelspecHandle, h_img: *const f32, n_mels: i32, width: i64, h_tga: *mut u8) -> i32;
        pub fn melspec_dequantize_tga_host(h: *mut MelspecHandle, h_tga: *const u8, tga_bytes: i64, h_img: *mut f32, capacity: i64) -> i32;
        // VAD over the mel image (src/vad.rs:251-338, 163-207)
        pub fn melspec_vad_default_settings(s: *mut VadSettings) -> i32;
        pub fn melspec_vad_host(
            h: *mut MelspecHandle,
            h_img: *const f32,
            n_mels: i32,
            width: i64,
            vs: *const VadSettings,
            h_smoothed: *mut u8,
            h_activity_opt: *mut i32,
        ) -> i32;
    }

    pub fn last_error() -> String {
        unsafe { CStr::from_ptr(melspec_last_error()).to_string_lossy().into_owned() }
    }
}
