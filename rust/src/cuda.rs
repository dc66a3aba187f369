//! Drop-in replacement for the reference's `src/cuda.rs` (wavey-ai/mel-spec): same public items
//! (`CudaError`, `CudaMelSpectrogram::{new, compute_mel_spectrogram, max_frames_per_batch}`, reference src/cuda.rs:10-155),
//! but everything below them is the melspec_b200 C ABI (`crate::ffi`) instead of cudart + cuFFT + `launch_mel_kernel`.
//!
//! SOURCE ONLY: this image has no cargo/rustc, so this file is not compiled or tested here; everything it calls
//! is exercised through the same C ABI from Python and C++ (`tests/`) — see INTEGRATION.md.
use crate::ffi;
use std::ffi::c_void;
use std::ptr;

/// Same shape as the reference's enum (src/cuda.rs:10-14): `Unavailable` carries a `&'static str`, so callers that match
/// on it keep compiling.  The library's own (dynamic) error text goes into `Runtime`.
#[derive(Debug)]
pub enum CudaError {
    Runtime(String),
    Unavailable(&'static str),
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

/// Status code of a failed `melspec_create` -> the reference's error kinds.  Constructor failures that mean "this machine
/// cannot run the backend" are `Unavailable` with a static text (the reference's own strings where one exists, src/cuda.rs:45-49,
/// 242-294); anything else (bad argument, CUDA runtime error) is `Runtime` with the library's message.
pub(crate) fn create_error(rc: i32) -> CudaError {
    match rc {
        ffi::ERR_NO_DEVICE => CudaError::Unavailable("no CUDA device of compute capability 10.0 (melspec_b200 has no CPU fallback)"),
        ffi::ERR_INVALID_CONFIG => CudaError::Unavailable("fft_size, hop_size, and n_mels must be non-zero"),
        ffi::ERR_UNSUPPORTED => CudaError::Unavailable("configuration not supported by this build of libmelspec_b200"),
        _ => CudaError::Runtime(ffi::last_error()),
    }
}

pub struct CudaMelSpectrogram {
    handle: *mut ffi::MelspecHandle, // raw pointer => !Send + !Sync, like the reference struct (src/cuda.rs:27-36)
    n_mels: usize,
}

impl CudaMelSpectrogram {
    /// Reference signature (src/cuda.rs:39-44).  Uses device 0; `new_on_device` picks another GPU of the box (one handle per
    /// GPU for the batch-sharded configuration).
    pub fn new(fft_size: usize, hop_size: usize, sampling_rate: f64, n_mels: usize) -> Result<Self, CudaError> {
        Self::new_on_device(fft_size, hop_size, sampling_rate, n_mels, 0)
    }

    pub fn new_on_device(fft_size: usize, hop_size: usize, sampling_rate: f64, n_mels: usize, device: i32) -> Result<Self, CudaError> {
        if fft_size == 0 || hop_size == 0 || n_mels == 0 {
            return Err(CudaError::Unavailable("fft_size, hop_size, and n_mels must be non-zero")); // src/cuda.rs:45-49
        }
        let mut cfg = ffi::MelspecConfig::default();
        unsafe { ffi::melspec_default_config(ffi::FRONTEND_WHISPER, &mut cfg) };
        cfg.fft_size = fft_size as i32;
        cfg.hop_size = hop_size as i32;
        cfg.n_mels = n_mels as i32;
        cfg.sampling_rate = sampling_rate;
        let mut handle = ptr::null_mut();
        let rc = unsafe { ffi::melspec_create(&cfg, device, &mut handle) };
        if rc != ffi::OK {
            return Err(create_error(rc));
        }
        Ok(Self { handle, n_mels })
    }

    pub fn max_frames_per_batch(&self) -> usize {
        unsafe { ffi::melspec_max_frames_per_batch(self.handle) as usize }
    }

    pub(crate) fn raw(&self) -> *mut ffi::MelspecHandle {
        self.handle
    }

    /// `&[f32]` -> `[frame][mel]`, the reference's signature (src/cuda.rs:88-101).
    pub fn compute_mel_spectrogram(&mut self, samples: &[f32]) -> Result<Vec<Vec<f32>>, CudaError> {
        let frames = unsafe { ffi::melspec_num_frames(self.handle, samples.len() as i64) } as usize;
        if frames == 0 {
            return Ok(Vec::new()); // src/cuda.rs:91-93
        }
        let mut flat = vec![0.0f32; frames * self.n_mels];
        let rc = unsafe {
            ffi::melspec_compute_host(self.handle, samples.as_ptr(), 1, samples.len() as i64, samples.len() as i64, flat.as_mut_ptr(),
                                      ffi::LAYOUT_FRAME_MAJOR, ptr::null_mut())
        };
        if rc != ffi::OK {
            return Err(CudaError::Runtime(ffi::last_error()));
        }
        Ok(flat.chunks(self.n_mels).map(|row| row.to_vec()).collect())
    }

    /// The same call for 16-bit PCM (half the bytes over PCIe; `x / 32768` on the device, bit-identical to the f32 call on the
    /// converted samples).
    pub fn compute_mel_spectrogram_i16(&mut self, samples: &[i16]) -> Result<Vec<Vec<f32>>, CudaError> {
        let frames = unsafe { ffi::melspec_num_frames(self.handle, samples.len() as i64) } as usize;
        if frames == 0 {
            return Ok(Vec::new());
        }
        let mut flat = vec![0.0f32; frames * self.n_mels];
        let rc = unsafe {
            ffi::melspec_compute_host_i16(self.handle, samples.as_ptr(), 1, samples.len() as i64, samples.len() as i64, flat.as_mut_ptr(),
                                          ffi::LAYOUT_FRAME_MAJOR, ptr::null_mut())
        };
        if rc != ffi::OK {
            return Err(CudaError::Runtime(ffi::last_error()));
        }
        Ok(flat.chunks(self.n_mels).map(|row| row.to_vec()).collect())
    }

    /// Device-resident batch entry (no host copies): `d_pcm` / `d_out` are CUDA device pointers.
    ///
    /// # Safety
    /// The pointers must be valid device allocations of the sizes described in `include/melspec_b200.h`.
    pub unsafe fn compute_device(&mut self, d_pcm: *const f32, n_clips: usize, clip_stride: usize, n_samples: usize, d_out: *mut f32,
                                 stream: *mut c_void) -> Result<(), CudaError> {
        let rc = ffi::melspec_compute_device(self.handle, d_pcm, n_clips as i64, clip_stride as i64, n_samples as i64, ptr::null(), d_out, 0,
                                             ffi::LAYOUT_FRAME_MAJOR, stream);
        if rc != ffi::OK {
            return Err(CudaError::Runtime(ffi::last_error()));
        }
        Ok(())
    }

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
            ffi::melspec_mel_tga_host(self.handle, samples.as_ptr(), samples.len() as i64, min_width as i64, out.as_mut_ptr(), size, &mut w,
                                      ptr::null_mut())
        };
        if rc != ffi::OK {
            return Err(CudaError::Runtime(ffi::last_error()));
        }
        Ok((out, w as usize))
    }

    /// `tga_8bit` (reference src/quant.rs:100-137): images wider than a TARGA can hold are cut into strides of at most 65 534
    /// columns (`chunk_frames_into_strides`), one TGA per stride.  `image` is row-major (n_mels, width).
    pub fn tga_8bit(&mut self, image: &[f32], n_mels: usize) -> Result<Vec<Vec<u8>>, CudaError> {
        const STRIDE: usize = u16::MAX as usize; // src/quant.rs:31
        let width = image.len() / n_mels;
        let mut out = Vec::new();
        let mut c0 = 0usize;
        while c0 < width {
            let w = STRIDE.min(width - c0);
            let mut strip = vec![0.0f32; n_mels * w];
            for m in 0..n_mels {
                strip[m * w..(m + 1) * w].copy_from_slice(&image[m * width + c0..m * width + c0 + w]);
            }
            let size = unsafe { ffi::melspec_tga_size(n_mels as i32, w as i64) } as usize;
            let mut tga = vec![0u8; size];
            let rc = unsafe { ffi::melspec_quantize_tga_host(self.handle, strip.as_ptr(), n_mels as i32, w as i64, tga.as_mut_ptr()) };
            if rc != ffi::OK {
                return Err(CudaError::Runtime(ffi::last_error()));
            }
            out.push(tga);
            c0 += w;
        }
        Ok(out)
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
        if rc != ffi::OK {
            return Err(CudaError::Runtime(ffi::last_error()));
        }
        Ok(mask.into_iter().map(|b| b != 0).collect())
    }
}

impl Drop for CudaMelSpectrogram {
    fn drop(&mut self) {
        unsafe { ffi::melspec_destroy(self.handle) }; // replaces src/cuda.rs:142-148
    }
}
