// melspec_b200.hpp — C++17 host-side mirror of the wavey-ai/mel-spec prelude over the C ABI of melspec_b200.h.
//
// The reference's host language is Rust; this build image has no cargo/rustc, so the compiled-language host side above
// the C ABI is this header (the Rust shim a maintainer would drop into the crate is kept, source-only, under rust/).
// Same names, argument meaning and error behaviour as the reference items cited on each declaration; everything numeric
// happens in libmelspec_b200.so (there is no CPU path here).  Header-only; link with -lmelspec_b200.
//
//   MelConfig                       src/config.rs:1-34
//   CudaError                       src/cuda.rs:10-25      (Runtime / Unavailable; thrown instead of returned)
//   CudaMelSpectrogram              src/cuda.rs:27-155
//   Spectrogram::compute_mel_spectrogram   batch semantics of src/stft.rs:119-138 (GPU-backed)
//   FbankConfig / Fbank             src/fbank.rs:25-82, 84-250
//   BatchLogMelConfig / BatchLogMelSpectrogram   src/mel.rs:171-418
//   RingBuffer                      src/rb.rs:12-122
//   interleave_frames               src/mel.rs:480-544
//   QuantizationRange / quantize / dequantize / tga_8bit_data / parse_tga_8bit   src/quant.rs:5-165
//   DetectionSettings / EdgeInfo / vad_boundaries / vad_on / VoiceActivity      src/vad.rs:5-338
#ifndef MELSPEC_B200_HPP_
#define MELSPEC_B200_HPP_

#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "melspec_b200.h"

namespace mel_spec {

// reference src/cuda.rs:10-25
class CudaError : public std::runtime_error {
public:
    enum class Kind { Runtime, Unavailable };
    CudaError(Kind k, const std::string& msg)
        : std::runtime_error((k == Kind::Unavailable ? "CUDA unavailable: " : "CUDA error: ") + msg), kind(k) {}
    Kind kind;
};

namespace detail {
inline void check(int32_t rc, bool constructing = false) {
    if (rc == MELSPEC_OK) return;
    const std::string msg = melspec_last_error();
    if (constructing && (rc == MELSPEC_ERR_NO_DEVICE || rc == MELSPEC_ERR_INVALID_CONFIG || rc == MELSPEC_ERR_UNSUPPORTED))
        throw CudaError(CudaError::Kind::Unavailable, msg);   // src/cuda.rs:45-49, 242-294
    if (rc == MELSPEC_ERR_INVALID_ARG) throw std::invalid_argument(msg);
    throw CudaError(CudaError::Kind::Runtime, msg);
}

// RAII over melspec_handle (the reference's Drop, src/cuda.rs:142-148); move-only like the !Send/!Sync Rust struct
class Handle {
public:
    Handle() = default;
    Handle(const melspec_config& cfg, int device) { check(melspec_create(&cfg, device, &h_), true); }
    Handle(const Handle&) = delete;
    Handle& operator=(const Handle&) = delete;
    Handle(Handle&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    Handle& operator=(Handle&& o) noexcept {
        if (this != &o) { reset(); h_ = o.h_; o.h_ = nullptr; }
        return *this;
    }
    ~Handle() { reset(); }
    void reset() {
        if (h_) melspec_destroy(h_);
        h_ = nullptr;
    }
    melspec_handle* get() const { return h_; }

private:
    melspec_handle* h_ = nullptr;
};
}  // namespace detail

// reference src/config.rs:1-34
struct MelConfig {
    size_t fft_size, hop_size, n_mels;
    double sampling_rate;
    MelConfig(size_t fft, size_t hop, size_t mels, double sr) : fft_size(fft), hop_size(hop), n_mels(mels), sampling_rate(sr) {}
};

// reference src/quant.rs:5-9
struct QuantizationRange {
    float min, max;
};

// reference src/vad.rs:5-22
struct DetectionSettings {
    double min_energy = 0.98;
    size_t min_y = 11, min_x = 5, min_mel = 2;
};

// reference src/vad.rs:488-520 (gradient_positions is left empty by the reference's vad_boundaries)
struct EdgeInfo {
    std::vector<size_t> non_intersected_columns, intersected_columns;
    const std::vector<size_t>& non_intersected() const { return non_intersected_columns; }
    const std::vector<size_t>& intersected() const { return intersected_columns; }
};

// reference src/vad.rs:226-249 (host logic over the device mask; quirk kept: the first column alone never fires)
inline bool vad_on(const EdgeInfo& e, size_t n) {
    const auto& c = e.intersected_columns;
    if (c.empty()) return false;
    size_t cnt = 1, prev = c[0];
    for (size_t i = 1; i < c.size(); ++i) {
        cnt = (c[i] == prev + 1) ? cnt + 1 : 1;
        if (cnt >= n) return true;
        prev = c[i];
    }
    return false;
}

// reference src/vad.rs:124-133
struct VoiceActivity {
    bool active;
    size_t frame_index, leading_active_columns, active_columns, window_columns;
    double confidence;
};

// reference src/cuda.rs:27-155 (+ the format / VAD steps that follow the path, on the same handle)
class CudaMelSpectrogram {
public:
    // CudaMelSpectrogram::new (src/cuda.rs:39-82)
    CudaMelSpectrogram(size_t fft_size, size_t hop_size, double sampling_rate, size_t n_mels, int device = 0) : n_mels_(n_mels) {
        if (fft_size == 0 || hop_size == 0 || n_mels == 0)   // src/cuda.rs:45-49
            throw CudaError(CudaError::Kind::Unavailable, "fft_size, hop_size, and n_mels must be non-zero");
        melspec_config cfg;
        melspec_default_config(MELSPEC_FRONTEND_WHISPER, &cfg);
        cfg.fft_size = (int32_t)fft_size; cfg.hop_size = (int32_t)hop_size; cfg.n_mels = (int32_t)n_mels;
        cfg.frame_length = (int32_t)fft_size; cfg.sampling_rate = sampling_rate;
        h_ = detail::Handle(cfg, device);
    }

    size_t max_frames_per_batch() const { return (size_t)melspec_max_frames_per_batch(h_.get()); }   // src/cuda.rs:84-86
    size_t num_frames(size_t n_samples) const { return (size_t)melspec_num_frames(h_.get(), (int64_t)n_samples); }
    size_t n_mels() const { return n_mels_; }
    melspec_handle* raw() const { return h_.get(); }

    // &[f32] -> Vec<Vec<f32>> [frame][mel] (src/cuda.rs:88-101); empty / too-short input => empty vector
    // opt-in 16-bit PCM (x / 32768 on the device; bit-identical to the f32 call on the converted samples)
    std::vector<std::vector<float>> compute_mel_spectrogram_i16(const std::vector<int16_t>& samples) {
        const int64_t frames = melspec_num_frames(h_.get(), (int64_t)samples.size());
        std::vector<std::vector<float>> out;
        if (frames <= 0) return out;
        std::vector<float> flat((size_t)frames * n_mels_);
        detail::check(melspec_compute_host_i16(h_.get(), samples.data(), 1, (int64_t)samples.size(), (int64_t)samples.size(), flat.data(),
                                               MELSPEC_LAYOUT_FRAME_MAJOR, nullptr));
        for (int64_t f = 0; f < frames; ++f) out.emplace_back(flat.begin() + f * (long)n_mels_, flat.begin() + (f + 1) * (long)n_mels_);
        return out;
    }
    std::vector<std::vector<float>> compute_mel_spectrogram(const std::vector<float>& samples) {
        const size_t f = num_frames(samples.size());
        std::vector<std::vector<float>> out;
        if (f == 0) return out;
        std::vector<float> flat(f * n_mels_);
        detail::check(melspec_compute_host(h_.get(), samples.data(), 1, (int64_t)samples.size(), (int64_t)samples.size(), flat.data(),
                                           MELSPEC_LAYOUT_FRAME_MAJOR, nullptr));
        out.reserve(f);
        for (size_t k = 0; k < f; ++k) out.emplace_back(flat.begin() + k * n_mels_, flat.begin() + (k + 1) * n_mels_);
        return out;
    }

    // device-resident batch entry, asynchronous on `stream` (a cudaStream_t)
    void compute_device(const float* d_pcm, size_t n_clips, size_t clip_stride, size_t n_samples, float* d_out, void* stream = nullptr,
                        int32_t layout = MELSPEC_LAYOUT_FRAME_MAJOR, const int32_t* d_lens = nullptr) {
        detail::check(melspec_compute_device(h_.get(), d_pcm, (int64_t)n_clips, (int64_t)clip_stride, (int64_t)n_samples, d_lens, d_out, 0,
                                             layout, stream));
    }

    // interleave_frames(frames, false, min_width) of the mel frames of `samples` (src/mel.rs:480-544): flat row-major
    // (n_mels, width) image, written in that layout by the fused kernel.  Also returns the TGA bytes if asked.
    std::vector<float> interleave_frames(const std::vector<float>& samples, size_t min_width = 0, std::vector<uint8_t>* tga = nullptr,
                                         size_t* width_out = nullptr) {
        if (min_width % 2) throw std::invalid_argument("min_width must be even");          // src/mel.rs:488
        const size_t f = num_frames(samples.size());
        if (f == 0) throw std::invalid_argument("frames is empty");                         // src/mel.rs:487
        const int64_t w = melspec_interleaved_width((int64_t)f, (int64_t)min_width);
        const int64_t size = melspec_tga_size((int32_t)n_mels_, w);
        if (size < 0) throw std::invalid_argument("width greater than TARGA max, use [`tga_8bit`]");
        std::vector<uint8_t> bytes((size_t)size);
        std::vector<float> img(n_mels_ * (size_t)w);
        int64_t wo = 0;
        detail::check(melspec_mel_tga_host(h_.get(), samples.data(), (int64_t)samples.size(), (int64_t)min_width, bytes.data(), size, &wo,
                                           img.data()));
        if (tga) *tga = std::move(bytes);
        if (width_out) *width_out = (size_t)wo;
        return img;
    }

    // PCM batch -> one TGA image per clip (interleave_frames + tga_8bit_data of every clip's frames, pipelined on the device).
    // `samples` holds n_clips rows of n_samples; returns n_clips * melspec_tga_size(n_mels, width) bytes, images back to back.
    std::vector<uint8_t> mel_tga_batch(const std::vector<float>& samples, size_t n_clips, size_t n_samples, size_t min_width = 0,
                                       size_t* width_out = nullptr) {
        if (min_width % 2) throw std::invalid_argument("min_width must be even");          // src/mel.rs:488
        if (samples.size() < n_clips * n_samples) throw std::invalid_argument("samples shorter than n_clips * n_samples");
        const size_t f = num_frames(n_samples);
        if (f == 0) throw std::invalid_argument("frames is empty");                         // src/mel.rs:487
        const int64_t w = melspec_interleaved_width((int64_t)f, (int64_t)min_width);
        const int64_t size = melspec_tga_size((int32_t)n_mels_, w);
        if (size < 0) throw std::invalid_argument("width greater than TARGA max, use [`tga_8bit`]");
        std::vector<uint8_t> bytes((size_t)size * n_clips);
        int64_t wo = 0;
        detail::check(melspec_mel_tga_host_batch(h_.get(), samples.data(), (int64_t)n_clips, (int64_t)n_samples, (int64_t)n_samples,
                                                 (int64_t)min_width, bytes.data(), 0, &wo));
        if (width_out) *width_out = (size_t)wo;
        return bytes;
    }

    // tga_8bit_data (src/quant.rs:38-64): row-major (n_mels, width) image -> TGA bytes
    std::vector<uint8_t> tga_8bit_data(const std::vector<float>& data, size_t n_mels) {
        if (n_mels == 0 || data.empty() || data.size() % n_mels) throw std::invalid_argument("data length must be a positive multiple of n_mels");
        const int64_t width = (int64_t)(data.size() / n_mels), size = melspec_tga_size((int32_t)n_mels, width);
        if (size < 0) throw std::invalid_argument("width greater than TARGA max, use [`tga_8bit`]");   // src/quant.rs:18-21
        std::vector<uint8_t> out((size_t)size);
        detail::check(melspec_quantize_tga_host(h_.get(), data.data(), (int32_t)n_mels, width, out.data()));
        return out;
    }

    // quantize (src/quant.rs:140-152)
    std::pair<std::vector<uint8_t>, QuantizationRange> quantize(const std::vector<float>& frame) {
        std::vector<uint8_t> tga = tga_8bit_data(frame, 1);
        QuantizationRange r;
        std::memcpy(&r.min, tga.data() + 18, 4);
        std::memcpy(&r.max, tga.data() + 22, 4);
        return {std::vector<uint8_t>(tga.begin() + 26, tga.end()), r};
    }

    // parse_tga_8bit (src/quant.rs:66-88)
    std::vector<float> parse_tga_8bit(const std::vector<uint8_t>& data) {
        if (data.size() < 26) throw std::runtime_error("failed to fill whole buffer");
        std::vector<float> out(data.size() - 26);
        detail::check(melspec_dequantize_tga_host(h_.get(), data.data(), (int64_t)data.size(), out.data(), (int64_t)out.size()));
        return out;
    }

    // vad_boundaries (src/vad.rs:251-338) on a row-major (n_mels, width) image
    EdgeInfo vad_boundaries(const std::vector<float>& image, size_t n_mels, const DetectionSettings& s = DetectionSettings()) {
        EdgeInfo e;
        const size_t width = n_mels ? image.size() / n_mels : 0;
        if (n_mels < 3 || width < 3) return e;                                              // src/vad.rs:264-266
        std::vector<uint8_t> mask(width - 2);
        const melspec_vad_settings vs{s.min_energy, (int32_t)s.min_y, (int32_t)s.min_x, (int32_t)s.min_mel};
        detail::check(melspec_vad_host(h_.get(), image.data(), (int32_t)n_mels, (int64_t)width, &vs, mask.data(), nullptr));
        for (size_t x = 0; x < mask.size(); ++x) (mask[x] ? e.intersected_columns : e.non_intersected_columns).push_back(x);
        return e;
    }

    // VoiceActivityDetector::add_activity (src/vad.rs:163-207) for every column of the image
    std::vector<VoiceActivity> vad_activities(const std::vector<float>& image, size_t n_mels, const DetectionSettings& s = DetectionSettings()) {
        std::vector<VoiceActivity> out;
        const size_t width = n_mels ? image.size() / n_mels : 0;
        if (width == 0) return out;
        std::vector<uint8_t> mask(width > 2 ? width - 2 : 1);
        std::vector<int32_t> act(3 * width);
        const melspec_vad_settings vs{s.min_energy, (int32_t)s.min_y, (int32_t)s.min_x, (int32_t)s.min_mel};
        detail::check(melspec_vad_host(h_.get(), image.data(), (int32_t)n_mels, (int64_t)width, &vs, mask.data(), act.data()));
        const size_t win = (n_mels >= 3 && s.min_x >= 3) ? s.min_x - 2 : 0;
        for (size_t i = 0; i < width; ++i) {
            if (act[3 * i] < 0) continue;
            out.push_back(VoiceActivity{act[3 * i] != 0, i, (size_t)act[3 * i + 1], (size_t)act[3 * i + 2], win,
                                        win ? (double)act[3 * i + 2] / (double)win : 0.0});
        }
        return out;
    }

private:
    detail::Handle h_;
    size_t n_mels_;
};

// What Spectrogram::add hands to MelSpectrogram::add.  The reference passes the complex FFT frame (src/stft.rs:82, src/mel.rs:26);
// here the chain is one fused kernel, so the token already holds the mel frame.
struct SpectrogramFrame {
    std::vector<float> mel;
    size_t fft_size = 0;
};

// reference src/stft.rs:10-138.  The static member is the batch entry (src/stft.rs:119-138; a handle per call like the reference
// re-plans per call); an instance is the streaming overlap-and-save entry with Spectrogram::add's own contract (src/stft.rs:48-86).
class Spectrogram {
public:
    static std::vector<std::vector<float>> compute_mel_spectrogram(const std::vector<float>& samples, size_t fft_size, size_t hop_size,
                                                                   size_t n_mels, double sampling_rate) {
        CudaMelSpectrogram m(fft_size, hop_size, sampling_rate, n_mels);
        return m.compute_mel_spectrogram(samples);
    }
    Spectrogram(size_t fft_size, size_t hop_size, size_t n_mels = 80, double sampling_rate = 16000.0, int device = 0)
        : fft_size_(fft_size), hop_size_(hop_size), n_mels_(n_mels), mel_(fft_size, hop_size, sampling_rate, n_mels, device) {
        detail::check(melspec_stream_create(mel_.raw(), (int64_t)hop_size, &s_), true);
    }
    Spectrogram(const Spectrogram&) = delete;
    Spectrogram& operator=(const Spectrogram&) = delete;
    ~Spectrogram() {
        if (s_) melspec_stream_destroy(s_);
    }
    // <= hop_size samples per call (more: std::invalid_argument, the reference asserts at src/stft.rs:53); a short chunk is
    // zero-padded to a whole hop; a frame once fft_size true samples have been seen and with every call after that
    std::optional<SpectrogramFrame> add(const std::vector<float>& frames) {
        if (frames.size() > hop_size_) throw std::invalid_argument("frames must be <= hop_size");
        SpectrogramFrame f;
        f.mel.resize(n_mels_);
        f.fft_size = fft_size_;
        int32_t emitted = 0;
        detail::check(melspec_stream_push_hop(s_, frames.data(), (int64_t)frames.size(), f.mel.data(), &emitted));
        if (!emitted) return std::nullopt;
        return f;
    }

private:
    size_t fft_size_, hop_size_, n_mels_;
    CudaMelSpectrogram mel_;
    melspec_stream* s_ = nullptr;
};

// reference src/mel.rs:13-32: add(fft_frame) -> n_mels values (the reference's (n_mels, 1) array)
class MelSpectrogram {
public:
    MelSpectrogram(size_t fft_size, double /*sampling_rate*/, size_t n_mels) : fft_size_(fft_size), n_mels_(n_mels) {}
    std::vector<float> add(const SpectrogramFrame& fft) const {
        if (fft.fft_size != fft_size_ || fft.mel.size() != n_mels_) throw std::invalid_argument("frame from a different configuration");
        return fft.mel;
    }

private:
    size_t fft_size_, n_mels_;
};

// reference src/fbank.rs:25-64 with its Default
struct FbankConfig {
    double sample_rate = 16000.0;
    size_t num_mel_bins = 80;
    double frame_length_ms = 25.0, frame_shift_ms = 10.0, dither = 0.0, energy_floor = 0.0;
    bool use_energy = false, use_log_fbank = true, use_power = true;
    double preemphasis = 0.97;
    bool apply_cmn = true;
    double low_freq = 20.0, high_freq = 0.0;
    size_t frame_length_samples() const { return (size_t)(sample_rate * frame_length_ms / 1000.0); }   // src/fbank.rs:68-70
    size_t frame_shift_samples() const { return (size_t)(sample_rate * frame_shift_ms / 1000.0); }     // src/fbank.rs:73-75
};

// reference src/fbank.rs:84-236: compute() -> row-major (T, num_mel_bins) f32
class Fbank {
public:
    explicit Fbank(const FbankConfig& c = FbankConfig(), int device = 0) : n_mels_(c.num_mel_bins) {
        melspec_config cfg;
        melspec_default_config(MELSPEC_FRONTEND_KALDI, &cfg);
        cfg.sampling_rate = c.sample_rate; cfg.n_mels = (int32_t)c.num_mel_bins;
        cfg.frame_length = (int32_t)c.frame_length_samples(); cfg.hop_size = (int32_t)c.frame_shift_samples();
        cfg.apply_cmn = c.apply_cmn; cfg.use_log_fbank = c.use_log_fbank; cfg.use_power = c.use_power;
        cfg.preemphasis = c.preemphasis; cfg.low_freq = c.low_freq; cfg.high_freq = c.high_freq; cfg.energy_floor = c.energy_floor;
        // `dither` and `use_energy` are carried by the reference's FbankConfig but never read by Fbank::compute
        // (src/fbank.rs:141-236): accepted and ignored here as well.
        h_ = detail::Handle(cfg, device);
    }
    // returns the flat (T, n_mels) matrix; `frames` receives T
    std::vector<float> compute(const std::vector<float>& samples, size_t* frames = nullptr) {
        const size_t t = (size_t)melspec_num_frames(h_.get(), (int64_t)samples.size());
        if (frames) *frames = t;
        std::vector<float> out(t * n_mels_);
        if (t)
            detail::check(melspec_compute_host(h_.get(), samples.data(), 1, (int64_t)samples.size(), (int64_t)samples.size(), out.data(),
                                               MELSPEC_LAYOUT_FRAME_MAJOR, nullptr));
        return out;
    }
    size_t n_mels() const { return n_mels_; }

private:
    detail::Handle h_;
    size_t n_mels_;
};

// reference src/mel.rs:171-208 with its Default
struct BatchLogMelConfig {
    size_t sample_rate = 16000, n_fft = 512, win_length = 400, hop_length = 160, n_mels = 80;
    double f_min = 0.0;
    std::optional<double> f_max;
    bool htk = false, norm = true;
    double preemphasis = 0.0;
    bool center = true;
    double log_zero_guard = 1.1920928955078125e-07;
    size_t pad_to = 0;
    bool normalize_per_feature = false;
};

// reference src/mel.rs:239-396: compute_flat -> feature-major (n_mels, padded_frames)
class BatchLogMelSpectrogram {
public:
    explicit BatchLogMelSpectrogram(const BatchLogMelConfig& c = BatchLogMelConfig(), int device = 0) : n_mels_(c.n_mels) {
        auto bad = [](const char* m) { throw std::invalid_argument(std::string("invalid log-mel config: ") + m); };   // src/mel.rs:656-683
        if (c.sample_rate == 0) bad("sample_rate must be > 0");
        if (c.n_fft == 0) bad("n_fft must be > 0");
        if (c.win_length == 0) bad("win_length must be > 0");
        if (c.win_length > c.n_fft) bad("win_length must be <= n_fft");
        if (c.hop_length == 0) bad("hop_length must be > 0");
        if (c.n_mels == 0) bad("n_mels must be > 0");
        if (!std::isfinite(c.log_zero_guard) || c.log_zero_guard <= 0.0) bad("log_zero_guard must be finite and > 0");
        melspec_config cfg;
        melspec_default_config(MELSPEC_FRONTEND_NEMO, &cfg);
        cfg.sampling_rate = (double)c.sample_rate; cfg.fft_size = (int32_t)c.n_fft; cfg.win_length = (int32_t)c.win_length;
        cfg.hop_size = (int32_t)c.hop_length; cfg.n_mels = (int32_t)c.n_mels; cfg.f_min = c.f_min; cfg.f_max = c.f_max.value_or(0.0);
        cfg.htk = c.htk; cfg.slaney_norm = c.norm; cfg.preemphasis = c.preemphasis; cfg.center = c.center;
        cfg.log_zero_guard = c.log_zero_guard; cfg.pad_to = (int32_t)c.pad_to; cfg.normalize_per_feature = c.normalize_per_feature;
        h_ = detail::Handle(cfg, device);
    }
    size_t padded_frames(size_t n_samples) const { return (size_t)melspec_padded_frames(h_.get(), (int64_t)n_samples); }
    std::vector<float> compute_flat(const std::vector<float>& samples, size_t* rows = nullptr, size_t* cols = nullptr) {
        const size_t c = padded_frames(samples.size());
        if (rows) *rows = n_mels_;
        if (cols) *cols = c;
        std::vector<float> out(n_mels_ * c, 0.0f);
        if (c)
            detail::check(melspec_compute_host(h_.get(), samples.data(), 1, (int64_t)samples.size(), (int64_t)samples.size(), out.data(),
                                               MELSPEC_LAYOUT_MEL_MAJOR, nullptr));
        return out;
    }

private:
    detail::Handle h_;
    size_t n_mels_;
};

// reference src/rb.rs:12-122: bounded sample FIFO (drops the oldest samples when full) -> whole hops go to the streaming
// C ABI -> maybe_mel() hands out one frame (n_mels values) at a time
class RingBuffer {
public:
    RingBuffer(const MelConfig& config, size_t capacity, int device = 0)
        : cfg_(config), capacity_(capacity), mel_(config.fft_size, config.hop_size, config.sampling_rate, config.n_mels, device) {
        detail::check(melspec_stream_create(mel_.raw(), (int64_t)config.hop_size, &s_), true);
        out_.resize(4 * config.n_mels);
    }
    RingBuffer(const RingBuffer&) = delete;
    RingBuffer& operator=(const RingBuffer&) = delete;
    ~RingBuffer() {
        if (s_) melspec_stream_destroy(s_);
    }
    void add_frame(const std::vector<float>& samples) {      // src/rb.rs:54-70
        for (float v : samples) add(v);
    }
    void add(float sample) {                                  // src/rb.rs:72-84
        if (fifo_.size() == capacity_ && capacity_ > 0) fifo_.pop_front();
        if (capacity_ > 0) fifo_.push_back(sample);
    }
    // src/rb.rs:86-121: one hop of queued samples -> at most one frame
    std::optional<std::vector<float>> maybe_mel() {
        if (fifo_.size() < cfg_.hop_size) return std::nullopt;
        std::vector<float> hop(fifo_.begin(), fifo_.begin() + (long)cfg_.hop_size);
        fifo_.erase(fifo_.begin(), fifo_.begin() + (long)cfg_.hop_size);
        int64_t emitted = 0;
        detail::check(melspec_stream_push(s_, hop.data(), (int64_t)hop.size(), out_.data(), 4, &emitted));
        if (emitted == 0) return std::nullopt;
        return std::vector<float>(out_.begin(), out_.begin() + (long)cfg_.n_mels);
    }

private:
    MelConfig cfg_;
    size_t capacity_;
    CudaMelSpectrogram mel_;
    melspec_stream* s_ = nullptr;
    std::deque<float> fifo_;
    std::vector<float> out_;
};

}  // namespace mel_spec
#endif  // MELSPEC_B200_HPP_
