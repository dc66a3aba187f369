//! melspec_b200 — the reference crate's prelude names (wavey-ai/mel-spec `src/prelude.rs:1-23`) backed by the B200 library.
//!
//! SOURCE ONLY in the build image (no cargo / rustc): see Cargo.toml.  Every item here is a thin owner of C-ABI handles
//! (`include/melspec_b200.h`); the arithmetic lives in `libmelspec_b200.so`.
pub mod batch;
pub mod config;
pub mod cuda;
pub mod fbank;
mod ffi;
pub mod rb;
pub mod stft;

pub mod prelude {
    pub use crate::batch::{BatchLogMelConfig, BatchLogMelError, BatchLogMelOutput, BatchLogMelSpectrogram};
    pub use crate::config::MelConfig;
    pub use crate::cuda::{CudaError, CudaMelSpectrogram};
    pub use crate::fbank::{Fbank, FbankConfig};
    pub use crate::rb::RingBuffer;
    pub use crate::stft::{MelSpectrogram, Spectrogram, SpectrogramFrame};
}
