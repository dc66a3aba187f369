//! `Fbank` / `FbankConfig` with the reference's signatures (src/fbank.rs:25-250), computed by libmelspec_b200.
use crate::cuda::{create_error, CudaError};
use crate::ffi;
use ndarray::Array2;
use std::ptr;

/// Field for field the reference's struct (src/fbank.rs:25-64); `use_energy` and `dither` are carried and ignored exactly
/// as `Fbank::compute` ignores them there (src/fbank.rs:141-236 never reads them).
#[derive(Clone, Debug)]
pub struct FbankConfig {
    pub sample_rate: f64,
    pub num_mel_bins: usize,
    pub frame_length_ms: f64,
    pub frame_shift_ms: f64,
    pub preemphasis: f64,
    pub low_freq: f64,
    pub high_freq: f64,
    pub energy_floor: f64,
    pub use_energy: bool,
    pub use_log_fbank: bool,
    pub use_power: bool,
    pub apply_cmn: bool,
    pub dither: f64,
}

impl Default for FbankConfig {
    fn default() -> Self {
        Self {
            sample_rate: 16000.0,
            num_mel_bins: 80,
            frame_length_ms: 25.0,
            frame_shift_ms: 10.0,
            preemphasis: 0.97,
            low_freq: 20.0,
            high_freq: 0.0,
            energy_floor: 0.0,
            use_energy: false,
            use_log_fbank: true,
            use_power: true,
            apply_cmn: true,
            dither: 0.0,
        }
    }
}

impl FbankConfig {
    pub fn frame_length_samples(&self) -> usize {
        (self.frame_length_ms / 1000.0 * self.sample_rate).round() as usize // src/fbank.rs:68-70
    }
    pub fn frame_shift_samples(&self) -> usize {
        (self.frame_shift_ms / 1000.0 * self.sample_rate).round() as usize // src/fbank.rs:73-75
    }
    pub fn fft_size(&self) -> usize {
        self.frame_length_samples().next_power_of_two() // src/fbank.rs:78-81
    }
}

pub struct Fbank {
    config: FbankConfig,
    handle: *mut ffi::MelspecHandle,
    dense: Array2<f64>,
}

impl Fbank {
    /// The reference's constructor is infallible (src/fbank.rs:94); here it needs a GPU, hence `try_new`.  `new` keeps the
    /// reference signature and panics with the `CudaError` text when no device is present.
    pub fn new(config: FbankConfig) -> Self {
        Self::try_new(config, 0).unwrap_or_else(|e| panic!("{e}"))
    }

    pub fn try_new(config: FbankConfig, device: i32) -> Result<Self, CudaError> {
        let mut cfg = ffi::MelspecConfig::default();
        unsafe { ffi::melspec_default_config(ffi::FRONTEND_KALDI, &mut cfg) };
        cfg.sampling_rate = config.sample_rate;
        cfg.n_mels = config.num_mel_bins as i32;
        cfg.frame_length = config.frame_length_samples() as i32;
        cfg.hop_size = config.frame_shift_samples() as i32;
        cfg.fft_size = config.fft_size() as i32;
        cfg.preemphasis = config.preemphasis;
        cfg.low_freq = config.low_freq;
        cfg.high_freq = config.high_freq;
        cfg.energy_floor = config.energy_floor;
        cfg.use_log_fbank = config.use_log_fbank as i32;
        cfg.use_power = config.use_power as i32;
        cfg.apply_cmn = config.apply_cmn as i32;
        let mut handle = ptr::null_mut();
        let rc = unsafe { ffi::melspec_create(&cfg, device, &mut handle) };
        if rc != ffi::OK {
            return Err(create_error(rc));
        }
        let bins = config.fft_size() / 2 + 1;
        let mut flat = vec![0.0f64; config.num_mel_bins * bins];
        unsafe { ffi::melspec_filterbank(handle, flat.as_mut_ptr(), flat.len() as i64) };
        let dense = Array2::from_shape_vec((config.num_mel_bins, bins), flat).expect("filterbank shape");
        Ok(Self { config, handle, dense })
    }

    /// `(T, num_mel_bins)` row-major f32, T = 1 + (len - frame_length) / frame_shift (src/fbank.rs:141-236).
    pub fn compute(&self, samples: &[f32]) -> Array2<f32> {
        let t = unsafe { ffi::melspec_num_frames(self.handle, samples.len() as i64) } as usize;
        let m = self.config.num_mel_bins;
        if t == 0 {
            return Array2::zeros((0, m)); // src/fbank.rs:147-151
        }
        let mut flat = vec![0.0f32; t * m];
        let rc = unsafe {
            ffi::melspec_compute_host(self.handle, samples.as_ptr(), 1, samples.len() as i64, samples.len() as i64, flat.as_mut_ptr(),
                                      ffi::LAYOUT_FRAME_MAJOR, ptr::null_mut())
        };
        assert!(rc == ffi::OK, "{}", CudaError::Runtime(ffi::last_error()));
        Array2::from_shape_vec((t, m), flat).expect("fbank output shape")
    }

    pub fn config(&self) -> &FbankConfig {
        &self.config
    }

    pub fn dense_filterbank(&self) -> &Array2<f64> {
        &self.dense
    }
}

impl Drop for Fbank {
    fn drop(&mut self) {
        unsafe { ffi::melspec_destroy(self.handle) };
    }
}
