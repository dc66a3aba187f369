//! `BatchLogMelSpectrogram` (NeMo / Parakeet frontend) with the reference's signatures (src/mel.rs:171-396).
use crate::ffi;
use ndarray::Array2;
use std::fmt;
use std::ptr;

#[derive(Clone, Debug)]
pub struct BatchLogMelConfig {
    pub sample_rate: usize,
    pub n_fft: usize,
    pub win_length: usize,
    pub hop_length: usize,
    pub n_mels: usize,
    pub f_min: f64,
    pub f_max: Option<f64>,
    pub htk: bool,
    pub norm: bool,
    pub preemphasis: f32,
    pub center: bool,
    pub log_zero_guard: f32,
    pub pad_to: usize,
    pub normalize_per_feature: bool,
}

impl Default for BatchLogMelConfig {
    fn default() -> Self {
        // src/mel.rs:189-208
        Self {
            sample_rate: 16000,
            n_fft: 512,
            win_length: 400,
            hop_length: 160,
            n_mels: 80,
            f_min: 0.0,
            f_max: None,
            htk: false,
            norm: true,
            preemphasis: 0.0,
            center: true,
            log_zero_guard: f32::EPSILON,
            pad_to: 0,
            normalize_per_feature: false,
        }
    }
}

#[derive(Debug)]
pub enum BatchLogMelError {
    InvalidConfig(&'static str),
    Shape(ndarray::ShapeError),
}

impl fmt::Display for BatchLogMelError {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        match self {
            Self::InvalidConfig(msg) => write!(f, "invalid log-mel config: {msg}"),
            Self::Shape(e) => write!(f, "invalid log-mel shape: {e}"),
        }
    }
}

impl std::error::Error for BatchLogMelError {}

impl From<ndarray::ShapeError> for BatchLogMelError {
    fn from(e: ndarray::ShapeError) -> Self {
        Self::Shape(e)
    }
}

pub struct BatchLogMelOutput {
    pub data: Vec<f32>,
    pub rows: usize,
    pub cols: usize,
}

pub struct BatchLogMelSpectrogram {
    config: BatchLogMelConfig,
    handle: *mut ffi::MelspecHandle,
}

impl BatchLogMelSpectrogram {
    pub fn new(config: BatchLogMelConfig) -> Result<Self, BatchLogMelError> {
        // validate_batch_config, src/mel.rs:656-683 (same order, same texts)
        if config.sample_rate == 0 {
            return Err(BatchLogMelError::InvalidConfig("sample_rate must be > 0"));
        }
        if config.n_fft == 0 {
            return Err(BatchLogMelError::InvalidConfig("n_fft must be > 0"));
        }
        if config.win_length == 0 {
            return Err(BatchLogMelError::InvalidConfig("win_length must be > 0"));
        }
        if config.win_length > config.n_fft {
            return Err(BatchLogMelError::InvalidConfig("win_length must be <= n_fft"));
        }
        if config.hop_length == 0 {
            return Err(BatchLogMelError::InvalidConfig("hop_length must be > 0"));
        }
        if config.n_mels == 0 {
            return Err(BatchLogMelError::InvalidConfig("n_mels must be > 0"));
        }
        if !config.log_zero_guard.is_finite() || config.log_zero_guard <= 0.0 {
            return Err(BatchLogMelError::InvalidConfig("log_zero_guard must be finite and > 0"));
        }
        let mut cfg = ffi::MelspecConfig::default();
        unsafe { ffi::melspec_default_config(ffi::FRONTEND_NEMO, &mut cfg) };
        cfg.sampling_rate = config.sample_rate as f64;
        cfg.fft_size = config.n_fft as i32;
        cfg.win_length = config.win_length as i32;
        cfg.frame_length = config.win_length as i32;
        cfg.hop_size = config.hop_length as i32;
        cfg.n_mels = config.n_mels as i32;
        cfg.f_min = config.f_min;
        cfg.f_max = config.f_max.unwrap_or(0.0);
        cfg.htk = config.htk as i32;
        cfg.slaney_norm = config.norm as i32;
        cfg.preemphasis = config.preemphasis as f64;
        cfg.center = config.center as i32;
        cfg.log_zero_guard = config.log_zero_guard as f64;
        cfg.pad_to = config.pad_to as i32;
        cfg.normalize_per_feature = config.normalize_per_feature as i32;
        let mut handle = ptr::null_mut();
        let rc = unsafe { ffi::melspec_create(&cfg, 0, &mut handle) };
        if rc != ffi::OK {
            return Err(BatchLogMelError::InvalidConfig("libmelspec_b200 could not create the frontend (no GPU, or unsupported size)"));
        }
        Ok(Self { config, handle })
    }

    pub fn config(&self) -> &BatchLogMelConfig {
        &self.config
    }

    /// `(n_mels, padded_frames)` feature-major f32 (src/mel.rs:299-303).
    pub fn compute(&self, samples: &[f32]) -> Result<Array2<f32>, BatchLogMelError> {
        let out = self.compute_flat(samples)?;
        Ok(Array2::from_shape_vec((out.rows, out.cols), out.data)?)
    }

    pub fn compute_flat(&self, samples: &[f32]) -> Result<BatchLogMelOutput, BatchLogMelError> {
        let cols = unsafe { ffi::melspec_padded_frames(self.handle, samples.len() as i64) } as usize;
        let rows = self.config.n_mels;
        let mut data = vec![0.0f32; rows * cols];
        if cols > 0 {
            let rc = unsafe {
                ffi::melspec_compute_host(self.handle, samples.as_ptr(), 1, samples.len() as i64, samples.len() as i64, data.as_mut_ptr(),
                                          ffi::LAYOUT_MEL_MAJOR, ptr::null_mut())
            };
            assert!(rc == ffi::OK, "{}", ffi::last_error());
        }
        Ok(BatchLogMelOutput { data, rows, cols })
    }
}

impl Drop for BatchLogMelSpectrogram {
    fn drop(&mut self) {
        unsafe { ffi::melspec_destroy(self.handle) };
    }
}
