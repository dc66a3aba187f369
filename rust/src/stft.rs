//! `Spectrogram` / `MelSpectrogram` with the reference's call shapes (src/stft.rs:10-138, src/mel.rs:13-32).
//!
//! The reference hands a complex FFT frame from `Spectrogram::add` to `MelSpectrogram::add`.  Here STFT, projection, log and
//! normalisation are one fused kernel, so the token passed between the two already carries the mel frame; the call sequence
//! `if let Some(f) = spec.add(&pcm) { let mel = mel.add(&f); }` compiles unchanged.
use crate::cuda::{CudaError, CudaMelSpectrogram};
use crate::ffi;
use ndarray::Array2;
use std::ptr;

pub struct SpectrogramFrame {
    mel: Vec<f32>,
    fft_size: usize,
}

pub struct Spectrogram {
    mel: CudaMelSpectrogram,
    stream: *mut ffi::MelspecStream,
    fft_size: usize,
    hop_size: usize,
    n_mels: usize,
    hop: Vec<f32>,
}

impl Spectrogram {
    /// Reference signature (src/stft.rs:25): Whisper's 80 mels at 16 kHz; `with_mel` states them explicitly.
    pub fn new(fft_size: usize, hop_size: usize) -> Self {
        Self::with_mel(fft_size, hop_size, 80, 16000.0).unwrap_or_else(|e| panic!("{e}"))
    }

    pub fn with_mel(fft_size: usize, hop_size: usize, n_mels: usize, sampling_rate: f64) -> Result<Self, CudaError> {
        let mel = CudaMelSpectrogram::new(fft_size, hop_size, sampling_rate, n_mels)?;
        let mut stream = ptr::null_mut();
        let rc = unsafe { ffi::melspec_stream_create(mel.raw(), hop_size as i64, &mut stream) };
        if rc != ffi::OK {
            return Err(CudaError::Runtime(ffi::last_error()));
        }
        Ok(Self { mel, stream, fft_size, hop_size, n_mels, hop: Vec::with_capacity(hop_size) })
    }

    /// src/stft.rs:48-86: `frames.len() <= hop_size` (asserted), a short chunk is zero-padded to a whole hop, `Some` once
    /// fft_size true samples have been seen and with every call after that.
    pub fn add(&mut self, frames: &[f32]) -> Option<SpectrogramFrame> {
        assert!(frames.len() <= self.hop_size, "frames must be <= hop_size");
        let mut out = vec![0.0f32; self.n_mels];
        let mut emitted = 0i32;
        let rc = unsafe { ffi::melspec_stream_push_hop(self.stream, frames.as_ptr(), frames.len() as i64, out.as_mut_ptr(), &mut emitted) };
        assert!(rc == ffi::OK, "{}", ffi::last_error());
        let _ = &self.hop;
        if emitted == 0 {
            None
        } else {
            Some(SpectrogramFrame { mel: out, fft_size: self.fft_size })
        }
    }

    /// Batch entry with the reference's signature (src/stft.rs:119-138).
    pub fn compute_mel_spectrogram_cpu(samples: &[f32], fft_size: usize, hop_size: usize, n_mels: usize, sampling_rate: f64) -> Vec<Vec<f32>> {
        let mut m = CudaMelSpectrogram::new(fft_size, hop_size, sampling_rate, n_mels).unwrap_or_else(|e| panic!("{e}"));
        m.compute_mel_spectrogram(samples).unwrap_or_else(|e| panic!("{e}"))
    }

    pub fn mel_handle(&self) -> &CudaMelSpectrogram {
        &self.mel
    }
}

impl Drop for Spectrogram {
    fn drop(&mut self) {
        unsafe { ffi::melspec_stream_destroy(self.stream) };
    }
}

pub struct MelSpectrogram {
    fft_size: usize,
    n_mels: usize,
}

impl MelSpectrogram {
    pub fn new(fft_size: usize, _sampling_rate: f64, n_mels: usize) -> Self {
        Self { fft_size, n_mels }
    }

    /// src/mel.rs:26-31: `(n_mels, 1)` f64.
    pub fn add(&mut self, fft: &SpectrogramFrame) -> Array2<f64> {
        assert!(fft.fft_size == self.fft_size && fft.mel.len() == self.n_mels, "frame from a different configuration");
        Array2::from_shape_vec((self.n_mels, 1), fft.mel.iter().map(|&v| v as f64).collect()).expect("mel output shape should match filterbank")
    }
}
