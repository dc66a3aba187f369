/*
 * melspec_b200 — C ABI of the B200-native log-mel frontend.
 *
 * This is the drop-in boundary for the STFT->mel hot path of wavey-ai/mel-spec: it replaces the private
 * `mod ffi` of the reference's CUDA backend (reference src/cuda.rs:161-480; the `unsafe extern "C"` block at
 * src/cuda.rs:185-220 with its 12 CUDA-runtime/cuFFT symbols and `launch_mel_kernel`, C side
 * src/cuda_kernels.cu:49-66) and sits behind
 *   - CudaMelSpectrogram::{new, compute_mel_spectrogram, max_frames_per_batch}   (src/cuda.rs:38-101,150-155)
 *   - Spectrogram::compute_mel_spectrogram_cpu's batch semantics                  (src/stft.rs:119-138)
 *   - Fbank::{new, compute}                                                       (src/fbank.rs:94-132,141-236)
 *   - RingBuffer::maybe_mel / Spectrogram::add streaming semantics                (src/rb.rs:86-121, src/stft.rs:48-86)
 *
 * Plain C: opaque handles, raw pointers and sizes, int32 status codes.  No torch / STL types cross this line.
 * All entry points are non-throwing.  A handle is single-threaded (the reference's struct is !Send/!Sync,
 * src/cuda.rs:27-36); different handles / devices may be used concurrently.  There is no CPU fallback: without
 * a CUDA device `melspec_create` fails with MELSPEC_ERR_NO_DEVICE (the reference's CudaError::Unavailable,
 * src/cuda.rs:10-25).
 */
#ifndef MELSPEC_B200_H_
#define MELSPEC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MELSPEC_B200_ABI_VERSION 2

/* ---- status codes (0 = success, like cudaError_t in src/cuda_kernels.cu:56-65) ---- */
enum {
    MELSPEC_OK = 0,
    MELSPEC_ERR_INVALID_CONFIG = 1, /* zero sizes (src/cuda.rs:45-49), unsupported fft size, ...           */
    MELSPEC_ERR_NO_DEVICE = 2,      /* CudaError::Unavailable                                              */
    MELSPEC_ERR_CUDA = 3,           /* CudaError::Runtime; text via melspec_last_error()                   */
    MELSPEC_ERR_INVALID_ARG = 4,    /* null pointers, negative sizes, capacity too small                   */
    MELSPEC_ERR_UNSUPPORTED = 5     /* well-formed request this build has no kernel for                    */
};

/* ---- which reference pipeline the handle implements ---- */
enum {
    MELSPEC_FRONTEND_WHISPER = 0, /* Hann -> FFT -> |X|^2 (bins < N/2) -> Slaney mel -> log10/1e-10 -> per-frame max-8, (x+4)/4
                                     src/stft.rs:89-169 + src/mel.rs:148-168,645-654                                         */
    MELSPEC_FRONTEND_KALDI = 1,   /* DC removal, pre-emphasis, Povey, zero-pad to 2^k, FFT, power, Kaldi mel (Hz triangles),
                                     ln(max(e, FLT_EPSILON)), optional CMN.  src/fbank.rs:141-236                            */
    MELSPEC_FRONTEND_NEMO = 2     /* BatchLogMelSpectrogram (NeMo/Parakeet style): whole-waveform pre-emphasis, centre zero
                                     padding, symmetric Hann of win_length centred in n_fft, power over bins 0..n_fft/2,
                                     Slaney mel, ln(e + guard), feature-major output padded to a multiple of pad_to, optional
                                     per-feature mean/std normalisation.  src/mel.rs:171-418, 656-756                        */
};

/* ---- output layouts ---- */
enum {
    MELSPEC_LAYOUT_FRAME_MAJOR = 0, /* out[clip][frame][mel]  == Vec<Vec<f32>> of src/stft.rs:119-138 and Array2 (T,80) of fbank */
    MELSPEC_LAYOUT_MEL_MAJOR = 1    /* out[clip][mel][frame]  == interleave_frames(.., false, 0) / rust_jfk_golden.npy layout    */
};

/*
 * POD configuration.  Mirrors MelConfig (src/config.rs:1-34) for the Whisper frontend and FbankConfig
 * (src/fbank.rs:25-64) for the Kaldi one.  Zero-initialise, then fill; `melspec_default_config` does it.
 */
typedef struct melspec_config {
    int32_t frontend;      /* MELSPEC_FRONTEND_*                                                               */
    int32_t fft_size;      /* Whisper / NeMo: N (any size).  Kaldi: ignored on input (next pow2 of frame_length). */
    int32_t hop_size;      /* samples between frames (Whisper hop / Kaldi frame shift)                         */
    int32_t n_mels;        /* 1..128                                                                           */
    double sampling_rate;  /* Hz                                                                               */
    /* Kaldi-only fields (FbankConfig).  Ignored for the Whisper frontend. */
    int32_t frame_length;  /* samples per frame before zero padding (400 at 25 ms / 16 kHz; any length >= 2)   */
    int32_t apply_cmn;     /* subtract per-mel mean over time (src/fbank.rs:226-233)                           */
    int32_t use_log_fbank; /* ln() of the floored energies                                                     */
    int32_t use_power;     /* 1: |X|^2, 0: |X| (src/fbank.rs:197-203)                                          */
    double preemphasis;    /* 0.97                                                                             */
    double low_freq;       /* 20 Hz                                                                            */
    double high_freq;      /* 0 => Nyquist                                                                     */
    double energy_floor;   /* <= 0 => FLT_EPSILON                                                              */
    /* NeMo-only fields (BatchLogMelConfig, src/mel.rs:171-208).  fft_size = n_fft, hop_size = hop_length,
       preemphasis is shared with the Kaldi block.  Ignored for the other frontends. */
    int32_t win_length;            /* 400                                                                      */
    int32_t center;                /* 1: pad n_fft/2 zeros on both sides, frames = len/hop + 1                 */
    int32_t pad_to;                /* output columns padded (with zeros) to a multiple of this; 0 = none       */
    int32_t normalize_per_feature; /* (x - mean) / (std + 1e-5) per mel row over the valid frames              */
    int32_t htk;                   /* mel scale: 0 Slaney, 1 HTK                                               */
    int32_t slaney_norm;           /* area-normalise the triangles                                             */
    double log_zero_guard;         /* added to the mel energy before ln(); must be finite and > 0              */
    double f_min;                  /* Hz                                                                       */
    double f_max;                  /* Hz, 0 => sampling_rate / 2                                               */
} melspec_config;

typedef struct melspec_handle melspec_handle;
typedef struct melspec_stream melspec_stream;

/* Fills `cfg` with the reference defaults: Whisper {400,160,80,16000} (README / src/cuda.rs:490-493) or
 * FbankConfig::default() (src/fbank.rs:46-64).  Returns MELSPEC_ERR_INVALID_ARG for an unknown frontend. */
int32_t melspec_default_config(int32_t frontend, melspec_config* cfg);

/* Builds the constant tables (window, twiddles, sparse banded filterbank) on `device` and returns a handle.
 * Replaces CudaMelSpectrogram::new (src/cuda.rs:39-82) / Fbank::new (src/fbank.rs:94-132).
 * Every size the reference accepts is served: the configurations of the reference's tests, goldens and benchmarks
 * (Whisper fft_size 400 at hop <= 256, Whisper fft_size 512 / hop 160, Kaldi 400-sample frames / hop 160 / power spectrum,
 * NeMo n_fft 512 / win_length 400 / hop 160) run on two specialised kernels, every other fft_size (<= 13652; odd sizes
 * <= 9009) / hop / frame length / sample rate on a general mixed-radix kernel with the same fusion and the same results
 * contract.  n_mels <= 128. */
int32_t melspec_create(const melspec_config* cfg, int32_t device, melspec_handle** out);

/* Replaces Drop (src/cuda.rs:142-148, 366-375).  NULL is a no-op. */
void melspec_destroy(melspec_handle* h);

/* Frames produced for a clip of `n_samples`: 0 if shorter than one frame, else (n - frame)/hop + 1
 * (src/stft.rs:153-157, src/fbank.rs:147-151); NeMo frontend: n/hop + 1 when centred (src/mel.rs:387-395). */
int64_t melspec_num_frames(const melspec_handle* h, int64_t n_samples);

/* Output columns per clip for the NeMo frontend: melspec_num_frames rounded up to a multiple of pad_to
 * (src/mel.rs:751-756); equals melspec_num_frames for the other frontends. */
int64_t melspec_padded_frames(const melspec_handle* h, int64_t n_samples);

/* The reference's batching knob (src/cuda.rs:150-155: min(8192, 64 MiB / bytes-per-frame)); kept for API
 * parity.  This implementation has no such limit on the device path; the value is what the reference returns. */
int32_t melspec_max_frames_per_batch(const melspec_handle* h);

int32_t melspec_n_mels(const melspec_handle* h);
int32_t melspec_fft_size(const melspec_handle* h);
int32_t melspec_hop_size(const melspec_handle* h);

/* Device-free: the dense filterbank a handle created from `cfg` would project with, row-major
 * (n_mels, fft_size/2+1) f64 — `mel()` of src/mel.rs:547-589 for the Whisper frontend, `kaldi_mel_filterbank`
 * (src/fbank.rs:253-301) for the Kaldi one.  `capacity` in doubles.  Usable without a GPU (host logic tests). */
int32_t melspec_build_filterbank(const melspec_config* cfg, double* out, int64_t capacity);

/* Device-free frame count for `cfg` (same rule as melspec_num_frames). */
int64_t melspec_num_frames_cfg(const melspec_config* cfg, int64_t n_samples);

/* Dense filterbank the handle projects with, row-major (n_mels, fft_size/2+1) f64 — `mel()` of src/mel.rs:547-589
 * or `Fbank::dense_filterbank` (src/fbank.rs:243-245).  `capacity` in doubles. */
int32_t melspec_filterbank(const melspec_handle* h, double* out, int64_t capacity);

/*
 * The hot path.  Device-resident PCM in, device-resident features out, asynchronous on `stream`
 * (a cudaStream_t passed as void*; NULL = legacy default stream).  Replaces cufftExecZ2Z + launch_mel_kernel +
 * the host-side windowing and norm_mel_vec of src/cuda.rs:88-139.
 *
 *   d_pcm            n_clips rows of f32 samples, row r at d_pcm + r*clip_stride
 *   n_samples        samples per row (rows shorter than that: pass d_lens)
 *   d_lens           optional DEVICE array of n_clips int32 valid lengths (<= n_samples); NULL = all n_samples
 *   d_out            frame-major: [n_clips][F][n_mels], mel-major: [n_clips][n_mels][F], F = melspec_num_frames(n_samples)
 *                    (NeMo frontend: mel-major [n_clips][n_mels][melspec_padded_frames(n_samples)], padding columns zeroed);
 *                    clip r at d_out + r*out_clip_stride (floats; 0 = dense F*n_mels).  Frames past a short clip's
 *                    own frame count are left untouched (Whisper, Kaldi) or zero like the pad_to columns (NeMo, where
 *                    per-feature normalisation then runs over each clip's own valid frames; ragged NeMo batches run on
 *                    the general kernel).
 * Fast path (TMA bulk copies) needs 16-byte aligned d_pcm/d_out, clip_stride % 4 == 0 and n_samples % 4 == 0;
 * otherwise a slower cooperative-copy path of the same kernel is used (results identical).
 */
int32_t melspec_compute_device(melspec_handle* h, const float* d_pcm, int64_t n_clips, int64_t clip_stride,
                               int64_t n_samples, const int32_t* d_lens, float* d_out, int64_t out_clip_stride,
                               int32_t layout, void* stream);

/*
 * Host-buffer convenience with the reference's shape: &[f32] -> Vec<Vec<f32>> ([frame][mel], src/cuda.rs:88-101)
 * or Fbank::compute's (T, n_mels) (src/fbank.rs:141-236), for a batch of equally long clips.
 * H2D + kernel + D2H, pipelined over internal streams and pinned staging; blocks until h_out is complete
 * (the reference synchronises per batch, src/cuda.rs:129).  n_samples < frame length => 0 frames, MELSPEC_OK
 * (src/cuda.rs:91-93).  `frames_out` (optional) receives F.  Batches are pipelined clip-chunk by clip-chunk; a single long
 * Whisper clip (>= 16 MB of PCM) is cut along time into 8 MB pieces of whole warp tiles and pipelined the same way (result
 * bit-identical to one launch): one hour of 16 kHz audio returns in 4.6 ms.
 */
int32_t melspec_compute_host(melspec_handle* h, const float* h_pcm, int64_t n_clips, int64_t clip_stride,
                             int64_t n_samples, float* h_out, int32_t layout, int64_t* frames_out);

/*
 * The same call for 16-bit PCM (opt-in; the reference takes &[f32] only, its examples convert `sample as f32 / 32768.0`,
 * examples/vad_ten_eval/src/main.rs:298-299).  Samples cross PCIe as int16 -- half the host-to-device bytes of the f32
 * call, which is what bounds it end to end -- and become x / 32768 (exact in f32) on the device, so the result is
 * bit-identical to melspec_compute_host on the converted samples.  melspec_convert_i16_device is that conversion on its own
 * (device pointers, asynchronous on `stream`): d_out[r * out_stride + i] = d_in[r * in_stride + i] / 32768.
 */
int32_t melspec_compute_host_i16(melspec_handle* h, const int16_t* h_pcm, int64_t n_clips, int64_t clip_stride,
                                 int64_t n_samples, float* h_out, int32_t layout, int64_t* frames_out);
int32_t melspec_convert_i16_device(melspec_handle* h, const int16_t* d_in, int64_t n_rows, int64_t in_stride, int64_t n_samples,
                                   float* d_out, int64_t out_stride, void* stream);

/*
 * Streaming (overlap-and-save) front end with RingBuffer/Spectrogram::add semantics (src/rb.rs:86-121,
 * src/stft.rs:48-86) when fed whole hops: frame k covers stream samples [c + k*hop, c + k*hop + N),
 * c = ceil(N/hop)*hop - N; a trailing partial hop stays buffered and is never emitted.
 * `push` copies host samples to the device on a side stream, runs the fused kernel on the new frames only
 * (the last N-hop samples stay resident on the device), and returns the new frames frame-major in h_out.
 */
int32_t melspec_stream_create(melspec_handle* h, int64_t max_chunk_samples, melspec_stream** out);
int32_t melspec_stream_push(melspec_stream* s, const float* h_samples, int64_t n, float* h_out,
                            int64_t out_capacity_frames, int64_t* frames_emitted);
int32_t melspec_stream_reset(melspec_stream* s);
void melspec_stream_destroy(melspec_stream* s);

/*
 * Spectrogram::add's own contract (src/stft.rs:48-86), for callers that feed short chunks: every call takes n <= hop_size
 * samples (more: MELSPEC_ERR_INVALID_ARG, the reference asserts at stft.rs:53), pads a short chunk with zeros to a whole hop
 * (stft.rs:56-59), advances the overlap buffer by one hop, and counts only the n true samples (`idx`, stft.rs:64).  A frame
 * comes back once idx >= fft_size -- from then on with EVERY call (stft.rs:66) -- as the Whisper mel frame
 * MelSpectrogram::add would produce from the FFT frame the reference returns (src/mel.rs:26-31).  `*emitted` = 1 and n_mels
 * floats in h_out_frame, or 0 (the reference's None).  Same stream object as melspec_stream_push (do not mix the two on one
 * stream); melspec_stream_reset clears idx.  The stream must have been created with max_chunk_samples >= hop_size.
 */
int32_t melspec_stream_push_hop(melspec_stream* s, const float* h_samples, int64_t n, float* h_out_frame, int32_t* emitted);

/*
 * ---- output formats: the step after the path (reference src/mel.rs:480-544, src/quant.rs:38-165) ----
 *
 * interleave_frames(frames, major_column_order = false, min_width): row-major (n_mels, W) f32 image of the Whisper mel
 * frames, W = melspec_interleaved_width(F, min_width): an all-zero frame is appended when min_width > 0 and F is odd,
 * then zero columns up to min_width (src/mel.rs:497-516).  The fused kernel writes this layout directly
 * (d_out[clip][mel][W], clip r at d_out + r*out_clip_stride, 0 = dense); no separate transpose pass exists.
 * Errors like the reference's asserts: odd min_width, no frame at all (MELSPEC_ERR_INVALID_ARG).
 */
int64_t melspec_interleaved_width(int64_t n_frames, int64_t min_width);   /* -1: invalid arguments */
int32_t melspec_compute_interleaved_device(melspec_handle* h, const float* d_pcm, int64_t n_clips, int64_t clip_stride,
                                           int64_t n_samples, int64_t min_width, float* d_out, int64_t out_clip_stride,
                                           void* stream);

/*
 * tga_8bit_data / quantize (src/quant.rs:38-64,140-152) and parse_tga_8bit / dequantize (src/quant.rs:66-88,155-165) on the
 * device: 18-byte TGA header (type 3, 8 bpp, width/height little-endian u16) + 8-byte ID field holding f32 min and max of
 * the image + n_mels*width bytes ((v - min) * (255 / (max - min))).round().clamp(0, 255).  Bytes are bit-exact with the
 * reference for identical f32 input.  One image per clip: image r at d_img + r*img_stride floats (0 = dense), TGA r at
 * d_tga + r*tga_stride bytes (0 = melspec_tga_size).  The quantiser keeps its min / max partials in one workspace per handle: calls on
 * the same handle must be issued on one stream (or otherwise serialised); use one handle per stream for concurrent quantisation.  width and n_mels must fit the header's u16 fields (<= 65535: the stride
 * tga_8bit cuts wider images into, src/quant.rs:29-36, 100-137; save_tga_8bit additionally asserts width < 65535, src/quant.rs:17-21).
 */
int64_t melspec_tga_size(int32_t n_mels, int64_t width);                 /* 26 + n_mels*width, -1 if it cannot be a TGA */
int32_t melspec_quantize_tga_device(melspec_handle* h, const float* d_img, int64_t n_imgs, int64_t img_stride, int32_t n_mels,
                                    int64_t width, uint8_t* d_tga, int64_t tga_stride, void* stream);
int32_t melspec_dequantize_tga_device(melspec_handle* h, const uint8_t* d_tga, int64_t n_imgs, int64_t tga_stride,
                                      int32_t n_mels, int64_t width, float* d_img, int64_t img_stride, void* stream);
/* host-buffer conveniences (blocking) */
int32_t melspec_quantize_tga_host(melspec_handle* h, const float* h_img, int32_t n_mels, int64_t width, uint8_t* h_tga);
int32_t melspec_dequantize_tga_host(melspec_handle* h, const uint8_t* h_tga, int64_t tga_bytes, float* h_img, int64_t capacity);
/* PCM -> Whisper mel -> interleave -> TGA bytes in one call (examples/mel_tga/src/main.rs:23-76 as one device pipeline);
 * h_img_opt (optional, n_mels*W floats) also returns the interleaved f32 image. */
int32_t melspec_mel_tga_host(melspec_handle* h, const float* h_pcm, int64_t n_samples, int64_t min_width, uint8_t* h_tga,
                             int64_t capacity, int64_t* width_out, float* h_img_opt);
/* The same for a batch of clips, pipelined like melspec_compute_host (H2D of chunk i+1, kernels of chunk i, D2H of chunk i-1
 * overlap): clip i's TGA image (melspec_tga_size(n_mels, W) bytes, W = melspec_interleaved_width(frames, min_width), returned in
 * *width_out) lands at h_tga + i * tga_stride (tga_stride 0 = images back to back).  One byte per mel value crosses PCIe on the way
 * back instead of four; the _i16 form takes 16-bit PCM (x / 32768 on the device, like melspec_compute_host_i16).  Bytes equal
 * tga_8bit_data (src/quant.rs:38-64) of interleave_frames (src/mel.rs:480-544) of the clip's own frames. */
int32_t melspec_mel_tga_host_batch(melspec_handle* h, const float* h_pcm, int64_t n_clips, int64_t clip_stride, int64_t n_samples,
                                   int64_t min_width, uint8_t* h_tga, int64_t tga_stride, int64_t* width_out);
int32_t melspec_mel_tga_host_batch_i16(melspec_handle* h, const int16_t* h_pcm, int64_t n_clips, int64_t clip_stride, int64_t n_samples,
                                       int64_t min_width, uint8_t* h_tga, int64_t tga_stride, int64_t* width_out);

/*
 * ---- VAD over the mel image: the consumer directly downstream (reference src/vad.rs) ----
 *
 * DetectionSettings (src/vad.rs:5-22).  melspec_vad_boundaries_device = vad_boundaries (src/vad.rs:251-338) for a batch of
 * row-major (n_mels, width) f32 images (what melspec_compute_interleaved_device writes, or a dequantised TGA): per column
 * the count of 3x3 Sobel gradients with gx^2+gy^2 >= min_energy^2 over mel rows [min_mel, n_mels-2) against min_y
 * (classify_columns_in_frame, src/vad.rs:373-415; sobel_gradient_sq, 472-486; f64 like the reference), then the +-4 majority
 * vote (smooth_mask, 343-360).  d_smoothed[img][x] = 1 <=> column x is in EdgeInfo::intersected(), x in [0, width-2);
 * d_raw (optional) receives the unsmoothed decisions.  Images with n_mels < 3 or width < 3 produce nothing (empty EdgeInfo).
 * melspec_vad_activity_device = VoiceActivityDetector::add_activity (src/vad.rs:163-207) for all frame indices at once from
 * d_raw: d_activity[img][i] = (active, leading_active_columns, active_columns) over the window of the last min_x frames,
 * (-1,-1,-1) for i < min_x-1 (the reference returns None); window_columns = max(min_x-2, 0) where defined.
 */
typedef struct melspec_vad_settings {
    double min_energy; /* 0.98 */
    int32_t min_y;     /* 11   */
    int32_t min_x;     /* 5    */
    int32_t min_mel;   /* 2    */
} melspec_vad_settings;
int32_t melspec_vad_default_settings(melspec_vad_settings* s);
int32_t melspec_vad_boundaries_device(melspec_handle* h, const float* d_img, int64_t n_imgs, int64_t img_stride, int32_t n_mels,
                                      int64_t width, const melspec_vad_settings* vs, uint8_t* d_raw, uint8_t* d_smoothed,
                                      int64_t mask_stride, void* stream);
int32_t melspec_vad_activity_device(melspec_handle* h, const uint8_t* d_raw, int64_t n_imgs, int64_t mask_stride, int32_t n_mels,
                                    int64_t width, const melspec_vad_settings* vs, int32_t* d_activity, int64_t activity_stride,
                                    void* stream);
int32_t melspec_vad_host(melspec_handle* h, const float* h_img, int32_t n_mels, int64_t width, const melspec_vad_settings* vs,
                         uint8_t* h_smoothed, int32_t* h_activity_opt);

/*
 * ---- optional: gather of the output shards over NVLink / NVSwitch (one process per GPU, batch-sharded clips) ----
 *
 * The hot path has no collective: every GPU computes its own clips.  When one rank needs the whole batch, the shards are
 * gathered with one ncclAllGather.  NCCL is not a link-time dependency: libnccl.so.2 is looked up at run time (the copy already
 * loaded in the process -- e.g. PyTorch's -- is reused), and these entries return MELSPEC_ERR_UNSUPPORTED when none is found.
 *   melspec_nccl_unique_id   rank 0 creates the 128-byte ncclUniqueId; the host program hands it to the other ranks (any channel)
 *   melspec_nccl_init        every rank joins (collective call); the communicator belongs to the handle
 *   melspec_gather_nccl      d_full[r*count .. (r+1)*count) = rank r's d_shard[0 .. count), equal `count` (floats) on all ranks,
 *                            asynchronous on `stream`; uneven batches: pad the shard to the largest count
 *   melspec_nccl_destroy     also done by melspec_destroy
 */
int32_t melspec_nccl_unique_id(uint8_t* id128);
int32_t melspec_nccl_init(melspec_handle* h, const uint8_t* id128, int32_t rank, int32_t world_size);
int32_t melspec_gather_nccl(melspec_handle* h, const float* d_shard, int64_t count, float* d_full, void* stream);
int32_t melspec_nccl_destroy(melspec_handle* h);

/* Number of kernel launches issued through this handle so far (bench.py's `gpu_launches`). */
int64_t melspec_launch_count(const melspec_handle* h);

/* Thread-local description of the last failure in this thread ("" if none).  Never NULL.  The pointer stays valid until the
 * next FAILING call of this library on the same thread (successful calls do not touch it, other threads have their own
 * string): copy it before calling again, as the Rust shim does (CStr -> String, rust/src/cuda.rs). */
const char* melspec_last_error(void);

int32_t melspec_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MELSPEC_B200_H_ */
