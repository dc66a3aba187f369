//! `RingBuffer` with the reference's signatures (src/rb.rs:12-122): samples are queued on the host (a bounded FIFO that drops
//! the oldest samples when full, like the reference's VecDeque build), whole hops go to the streaming C ABI, and `maybe_mel`
//! hands out one `(n_mels, 1)` frame at a time.
use crate::config::MelConfig;
use crate::cuda::CudaMelSpectrogram;
use crate::ffi;
use ndarray::Array2;
use std::collections::VecDeque;
use std::ptr;

pub struct RingBuffer {
    config: MelConfig,
    capacity: usize,
    fifo: VecDeque<f32>,
    mel: CudaMelSpectrogram,
    stream: *mut ffi::MelspecStream,
    hop: Vec<f32>,
    frame: Vec<f32>,
}

impl RingBuffer {
    pub fn new(config: MelConfig, capacity: usize) -> Self {
        let mel = CudaMelSpectrogram::new(config.fft_size(), config.hop_size(), config.sampling_rate(), config.n_mels())
            .unwrap_or_else(|e| panic!("{e}"));
        let mut stream = ptr::null_mut();
        let rc = unsafe { ffi::melspec_stream_create(mel.raw(), config.hop_size() as i64, &mut stream) };
        assert!(rc == ffi::OK, "{}", ffi::last_error());
        let (hop, n_mels) = (config.hop_size(), config.n_mels());
        Self { config, capacity, fifo: VecDeque::with_capacity(capacity), mel, stream, hop: vec![0.0; hop], frame: vec![0.0; n_mels] }
    }

    pub fn add_frame(&mut self, samples: &[f32]) {
        for &s in samples {
            self.add(s); // src/rb.rs:54-70
        }
    }

    pub fn add(&mut self, sample: f32) {
        if self.fifo.len() == self.capacity {
            self.fifo.pop_front(); // src/rb.rs:72-84: the oldest sample makes room
        }
        self.fifo.push_back(sample);
    }

    /// One hop of queued samples -> at most one frame (src/rb.rs:86-121): `None` until a whole hop is queued and until the
    /// stream has seen fft_size samples; a trailing partial hop is never emitted.
    pub fn maybe_mel(&mut self) -> Option<Array2<f64>> {
        let hop = self.config.hop_size();
        if self.fifo.len() < hop {
            return None;
        }
        for (dst, src) in self.hop.iter_mut().zip(self.fifo.drain(..hop)) {
            *dst = src;
        }
        let mut emitted = 0i64;
        let rc = unsafe { ffi::melspec_stream_push(self.stream, self.hop.as_ptr(), hop as i64, self.frame.as_mut_ptr(), 1, &mut emitted) };
        assert!(rc == ffi::OK, "{}", ffi::last_error());
        if emitted == 0 {
            return None;
        }
        let col: Vec<f64> = self.frame.iter().map(|&v| v as f64).collect();
        Some(Array2::from_shape_vec((self.config.n_mels(), 1), col).expect("mel frame shape"))
    }
}

impl Drop for RingBuffer {
    fn drop(&mut self) {
        unsafe { ffi::melspec_stream_destroy(self.stream) }; // before `mel` (field order drops `mel` after this body)
    }
}
