// melspec_b200 host side: the C ABI declared in include/melspec_b200.h.
//
// Builds the constant tables (window, twiddles, sparse banded filterbank -> per-lane projection program) in f64 on
// the host, owns the device copies, and launches the fused kernel of melspec_kernels.cuh.  No cuFFT, no torch.
// Reference interfaces replaced: src/cuda.rs:39-155 (CudaMelSpectrogram), src/cuda.rs:161-480 (mod ffi),
// src/fbank.rs:94-236 (Fbank), src/rb.rs:86-121 + src/stft.rs:48-86 (streaming).
#include "../../include/melspec_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <functional>
#include <vector>

#include "melspec_kernels.cuh"
#include "melspec_generic.cuh"
#include "melspec_generic2.cuh"

namespace {

thread_local std::string g_last_error;

int32_t fail(int32_t code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
int32_t fail_cuda(cudaError_t e, const char* what) {
    g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
    return MELSPEC_ERR_CUDA;
}
#define MS_CUDA(call)                                           \
    do {                                                        \
        cudaError_t e__ = (call);                               \
        if (e__ != cudaSuccess) return fail_cuda(e__, #call);   \
    } while (0)

// ------------------------------------------------------------------------------------------------ filterbanks (f64)
// Slaney / HTK mel scales and librosa-style triangles: the math of SURVEY Appendix A.1 / reference src/mel.rs:547-643.
double mel_from_hz(double f, bool htk) {
    if (htk) return 2595.0 * std::log10(1.0 + f / 700.0);
    const double f_sp = 200.0 / 3.0, brk = 1000.0, brk_mel = brk / f_sp, step = std::log(6.4) / 27.0;
    return f >= brk ? brk_mel + std::log(f / brk) / step : f / f_sp;
}
double hz_from_mel(double m, bool htk) {
    if (htk) return 700.0 * (std::pow(10.0, m / 2595.0) - 1.0);
    const double f_sp = 200.0 / 3.0, brk = 1000.0, brk_mel = brk / f_sp, step = std::log(6.4) / 27.0;
    return m >= brk_mel ? brk * std::exp(step * (m - brk_mel)) : f_sp * m;
}
void build_slaney(double sr, int n_fft, int n_mels, double f_min, double f_max, bool htk, bool norm, std::vector<double>& w) {
    const int nb = n_fft / 2 + 1;
    w.assign((size_t)n_mels * nb, 0.0);
    std::vector<double> edge(n_mels + 2);
    const double lo = mel_from_hz(f_min, htk), hi = mel_from_hz(f_max, htk);
    const double step = (hi - lo) / (double)(n_mels + 1);
    for (int i = 0; i < n_mels + 2; ++i) edge[i] = hz_from_mel(lo + step * (double)i, htk);
    for (int m = 0; m < n_mels; ++m) {
        const double up = edge[m + 1] - edge[m], dn = edge[m + 2] - edge[m + 1];
        const double area = norm ? 2.0 / (edge[m + 2] - edge[m]) : 1.0;
        for (int b = 0; b < nb; ++b) {
            const double f = (sr / (double)n_fft) * (double)b;
            const double rise = std::min(std::max((f - edge[m]) / up, 0.0), 1.0);
            const double fall = std::min(std::max((edge[m + 2] - f) / dn, 0.0), 1.0);
            w[(size_t)m * nb + b] = std::min(rise, fall) * area;
        }
    }
}
// Kaldi-mel edges, triangles evaluated in Hz, no area normalisation: reference src/fbank.rs:253-313.
void build_kaldi(double sr, int n_fft, int n_mels, double low, double high, std::vector<double>& w) {
    const int nb = n_fft / 2 + 1;
    w.assign((size_t)n_mels * nb, 0.0);
    const double ml = 1127.0 * std::log(1.0 + low / 700.0), mh = 1127.0 * std::log(1.0 + high / 700.0);
    std::vector<double> hz(n_mels + 2);
    for (int i = 0; i < n_mels + 2; ++i)
        hz[i] = 700.0 * (std::exp((ml + (mh - ml) * (double)i / (double)(n_mels + 1)) / 1127.0) - 1.0);
    for (int m = 0; m < n_mels; ++m) {
        const double l = hz[m], c = hz[m + 1], r = hz[m + 2];
        if (c <= l || r <= c) continue;
        for (int b = 0; b < nb; ++b) {
            const double f = (double)b * sr / (double)n_fft;
            if (f > l && f <= c) w[(size_t)m * nb + b] = (f - l) / (c - l);
            else if (f > c && f < r) w[(size_t)m * nb + b] = (r - f) / (r - c);
        }
    }
}

int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// Validated, normalised view of a melspec_config.
struct Resolved {
    int frontend, fft, hop, n_mels, frame_len;
    double sr, preemph, low, high, floor;
    int cmn, use_log, use_power;
    // NeMo frontend
    int center = 0, pad_to = 0, norm_feat = 0, htk = 0, slaney_norm = 1;
    double guard = 0.0;
};

int32_t resolve(const melspec_config* c, Resolved& r) {
    if (!c) return fail(MELSPEC_ERR_INVALID_ARG, "config is null");
    r.frontend = c->frontend;
    r.hop = c->hop_size;
    r.n_mels = c->n_mels;
    r.sr = c->sampling_rate;
    if (c->frontend == MELSPEC_FRONTEND_WHISPER) {
        r.fft = c->fft_size;
        r.frame_len = c->fft_size;
        if (r.fft <= 0 || r.hop <= 0 || r.n_mels <= 0)
            return fail(MELSPEC_ERR_INVALID_CONFIG, "fft_size, hop_size, and n_mels must be non-zero");   // src/cuda.rs:45-49
        r.preemph = 0; r.low = 0; r.high = r.sr / 2; r.floor = 1e-10; r.cmn = 0; r.use_log = 1; r.use_power = 1;
    } else if (c->frontend == MELSPEC_FRONTEND_KALDI) {
        r.frame_len = c->frame_length;
        if (r.frame_len <= 1 || r.hop <= 0 || r.n_mels <= 0)
            return fail(MELSPEC_ERR_INVALID_CONFIG, "frame_length, hop_size, and n_mels must be non-zero");
        r.fft = next_pow2(r.frame_len);
        r.preemph = c->preemphasis; r.low = c->low_freq; r.high = c->high_freq == 0.0 ? r.sr / 2.0 : c->high_freq;
        r.floor = c->energy_floor > 0.0 ? c->energy_floor : (double)1.1920928955078125e-07f;
        r.cmn = c->apply_cmn; r.use_log = c->use_log_fbank; r.use_power = c->use_power;
    } else if (c->frontend == MELSPEC_FRONTEND_NEMO) {   // validate_batch_config, src/mel.rs:656-683
        r.fft = c->fft_size;
        r.frame_len = c->win_length;
        if (r.fft <= 0) return fail(MELSPEC_ERR_INVALID_CONFIG, "n_fft must be > 0");
        if (r.frame_len <= 0) return fail(MELSPEC_ERR_INVALID_CONFIG, "win_length must be > 0");
        if (r.frame_len > r.fft) return fail(MELSPEC_ERR_INVALID_CONFIG, "win_length must be <= n_fft");
        if (r.hop <= 0) return fail(MELSPEC_ERR_INVALID_CONFIG, "hop_length must be > 0");
        if (r.n_mels <= 0) return fail(MELSPEC_ERR_INVALID_CONFIG, "n_mels must be > 0");
        if (!std::isfinite(c->log_zero_guard) || c->log_zero_guard <= 0.0)
            return fail(MELSPEC_ERR_INVALID_CONFIG, "log_zero_guard must be finite and > 0");
        r.preemph = c->preemphasis; r.low = c->f_min; r.high = c->f_max == 0.0 ? c->sampling_rate / 2.0 : c->f_max;
        r.floor = 0.0; r.cmn = 0; r.use_log = 1; r.use_power = 1;
        r.center = c->center; r.pad_to = c->pad_to < 0 ? 0 : c->pad_to; r.norm_feat = c->normalize_per_feature;
        r.htk = c->htk; r.slaney_norm = c->slaney_norm; r.guard = c->log_zero_guard;
    } else {
        return fail(MELSPEC_ERR_INVALID_CONFIG, "unknown frontend");
    }
    if (!(r.sr > 0.0)) return fail(MELSPEC_ERR_INVALID_CONFIG, "sampling_rate must be positive");
    if (r.n_mels > 32 * melspec::kMaxMpl) return fail(MELSPEC_ERR_INVALID_CONFIG, "n_mels must be <= 128");
    return MELSPEC_OK;
}

void build_filterbank(const Resolved& r, std::vector<double>& w) {
    if (r.frontend == MELSPEC_FRONTEND_WHISPER) build_slaney(r.sr, r.fft, r.n_mels, 0.0, r.sr / 2.0, false, true, w);
    else if (r.frontend == MELSPEC_FRONTEND_NEMO) build_slaney(r.sr, r.fft, r.n_mels, r.low, r.high, r.htk != 0, r.slaney_norm != 0, w);
    else build_kaldi(r.sr, r.fft, r.n_mels, r.low, r.high, w);
}

int64_t frames_for(const Resolved& r, int64_t n) {
    if (r.frontend == MELSPEC_FRONTEND_NEMO) {   // src/mel.rs:321-328, 387-395 (empty input => no columns)
        if (n <= 0) return 0;
        if (r.center) return n / r.hop + 1;
        return n < r.fft ? 0 : (n - r.fft) / r.hop + 1;
    }
    return n < r.frame_len ? 0 : (n - r.frame_len) / r.hop + 1;
}

int64_t padded_frames_for(const Resolved& r, int64_t n) {
    const int64_t f = frames_for(r, n);
    if (r.frontend != MELSPEC_FRONTEND_NEMO || r.pad_to <= 0) return f;
    return (f + r.pad_to - 1) / r.pad_to * r.pad_to;   // src/mel.rs:751-756
}

}  // namespace

// ------------------------------------------------------------------------------------------------ handle
struct melspec_handle {
    Resolved cfg;
    int device = 0;
    int num_sms = 0;
    int plan = 0;   // 400 or 512 (specialised kernels), 1 = the general plan (melspec_generic.cuh)
    std::vector<double> dense;   // (n_mels, fft/2+1)
    // general plan: twiddles W_N^k, window, CSR band table, radix schedule
    float2* d_gtw = nullptr;
    float* d_gwin = nullptr;
    int* d_gbands = nullptr;
    float* d_gweights = nullptr;
    std::vector<int> radices;
    int pair_nw = 0, pair_warps = 0;   // general plan, pair form: warps per CTA (0: not chosen yet, -1: does not fit) and per SM
    float* d_gweights_t = nullptr; // general plan, pair form: weights as [slot][entry][lane], zero padded
    int* d_gstarts_t = nullptr;    // ... and the first power row of every mel row's window
    int n_gweights_t = 0;
    int g_kmax[4] = {0, 0, 0, 0};  // entries per slot of that table
    // device tables
    float* d_window = nullptr;
    float4* d_twiddle = nullptr;
    float2* d_rot10 = nullptr;
    float2* d_proj = nullptr;
    int* d_meta = nullptr;
    int proj_ktot = 0;
    int kspec = 0;                 // compile-time projection schedule the table matches (1: Whisper-80/fft-400: 14,4,2)
    int ksched512 = 0;             // plan 512: index into melspec::ksched512 (0: none, counts are read from the table)
    int proj_wavefront_cost = 0;   // half-warp wavefronts per LDS.64 of the projection loop, summed over entries (ideal: 2 per entry)
    int mpl = 0;
    // host-path resources (lazily created)
    cudaStream_t streams[3] = {nullptr, nullptr, nullptr};
    float* d_slot_pcm[3] = {nullptr, nullptr, nullptr};
    float* d_slot_out[3] = {nullptr, nullptr, nullptr};
    int16_t* d_slot_i16[3] = {nullptr, nullptr, nullptr};   // int16 staging of melspec_compute_host_i16
    size_t slot_pcm_cap = 0, slot_out_cap = 0, slot_i16_cap = 0;
    float2* d_partials = nullptr;   // min/max partials of the TGA quantiser
    uint8_t* d_slot_tga[3] = {nullptr, nullptr, nullptr};   // melspec_mel_tga_host_batch: TGA bytes and min/max partials per pipeline slot
    float2* d_slot_part[3] = {nullptr, nullptr, nullptr};
    size_t slot_tga_cap = 0, slot_part_cap = 0;
    size_t partials_cap = 0;
    float* d_fmt_img = nullptr;     // staging of the host-buffer format entry points
    unsigned char* d_fmt_tga = nullptr;
    size_t fmt_img_cap = 0, fmt_tga_cap = 0;
    int64_t launches = 0;
    void* nccl_comm = nullptr;      // ncclComm_t of the optional output gather (melspec_nccl_init)
    int nccl_world = 0;
};

struct melspec_stream {
    melspec_handle* h = nullptr;
    int64_t max_chunk = 0;  // largest push the caller announced
    int64_t piece = 0;      // samples per internal pipeline piece
    float* d_buf[2] = {nullptr, nullptr};
    int cur = 0;
    int64_t cap = 0;        // samples per device buffer
    int64_t buffered = 0;   // valid samples at the start of d_buf[cur] (the carried tail)
    int64_t to_skip = 0;    // samples still to drop before the first frame (the stream offset c)
    uint64_t idx = 0;       // melspec_stream_push_hop: true samples seen so far (Spectrogram::add's idx, src/stft.rs:64)
    std::vector<float> hopbuf;   // melspec_stream_push_hop: the zero-padded hop
    float* h_pin_in[2] = {nullptr, nullptr};
    float* h_pin_out[2] = {nullptr, nullptr};
    float* d_out[2] = {nullptr, nullptr};
    int64_t out_cap_frames = 0;   // per piece
    cudaStream_t copy_stream = nullptr, compute_stream = nullptr;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr};    // staging slot consumed by its H2D copy
    cudaEvent_t ev_free[2] = {nullptr, nullptr};   // device buffer no longer read by a kernel / tail copy
    cudaEvent_t ev_d2h[2] = {nullptr, nullptr};    // output staging slot landed on the host
    bool used_free[2] = {false, false};
};

namespace {

// Tables of the general plan (melspec_generic.cuh): twiddles, window, CSR bands, radix schedule.
int32_t build_tables_generic(melspec_handle* h) {
    const Resolved& c = h->cfg;
    const int N = c.fft, nb = N / 2 + 1, L = c.frame_len;
    std::vector<float2> tw((size_t)N);
    for (int k = 0; k < N; ++k) {
        const double a = 2.0 * M_PI * (double)k / (double)N;
        tw[k] = make_float2((float)std::cos(a), (float)-std::sin(a));
    }
    std::vector<float> win((size_t)L, 0.f);
    for (int i = 0; i < L; ++i) {
        if (c.frontend == MELSPEC_FRONTEND_WHISPER)        // periodic Hann, src/stft.rs:141-145
            win[i] = (float)(0.5 * (1.0 - std::cos(2.0 * M_PI * (double)i / (double)N)));
        else if (c.frontend == MELSPEC_FRONTEND_NEMO)      // symmetric Hann of win_length (zeros if <= 1), src/mel.rs:708-719
            win[i] = L <= 1 ? 0.f : (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * (double)i / (double)(L - 1)));
        else                                               // Povey, src/fbank.rs:100-105
            win[i] = (float)std::pow(0.5 - 0.5 * std::cos(2.0 * M_PI * (double)i / (double)(L - 1)), 0.85);
    }
    // radix schedule: 8s, 4s, then a 2, then the odd prime factors in ascending order
    h->radices.clear();
    int n = (N % 2 == 0) ? N / 2 : N;   // even N: one complex N/2-point transform of the even/odd samples
    while (n % 8 == 0) { h->radices.push_back(8); n /= 8; }
    while (n % 4 == 0) { h->radices.push_back(4); n /= 4; }
    if (n % 2 == 0) { h->radices.push_back(2); n /= 2; }
    for (int f = 3; (long long)f * f <= n; f += 2)
        while (n % f == 0) { h->radices.push_back(f); n /= f; }
    if (n > 1) h->radices.push_back(n);
    if ((int)h->radices.size() > melspec::kMaxStages) return fail(MELSPEC_ERR_UNSUPPORTED, "fft_size has too many prime factors");
    // CSR bands: per mel row the run [first non-zero bin, last non-zero bin]; Whisper drops bins >= N/2 (src/mel.rs:158-162)
    const int last = c.frontend == MELSPEC_FRONTEND_WHISPER ? N / 2 - 1 : N / 2;
    std::vector<int> bands((size_t)3 * c.n_mels, 0);
    std::vector<float> wts;
    for (int m = 0; m < c.n_mels; ++m) {
        int b0 = -1, b1 = -1;
        for (int b = 0; b <= last; ++b)
            if (h->dense[(size_t)m * nb + b] != 0.0) { if (b0 < 0) b0 = b; b1 = b; }
        bands[3 * m] = b0 < 0 ? 0 : b0;
        bands[3 * m + 1] = b0 < 0 ? 0 : b1 - b0 + 1;
        bands[3 * m + 2] = (int)wts.size();
        for (int b = b0; b0 >= 0 && b <= b1; ++b) wts.push_back((float)h->dense[(size_t)m * nb + b]);
    }
    if (wts.empty()) wts.push_back(0.f);
    // pair form: the same weights as [slot][entry][lane], zero padded to the slot's longest window.  A mel row's window starts at or
    // up to 15 rows before its band (zero weights in front) so that the 16 lanes of a half-warp start at 16 different rows mod 16:
    // their 64-bit reads of the power pairs are then conflict-free (measured before: 5 wavefronts per read instead of 2).
    std::vector<float> wts_t;
    std::vector<int> starts((size_t)std::max(c.n_mels, 1), 0);
    for (int sl = 0; sl < melspec::kMaxMpl; ++sl) {
        int lead[32] = {0};
        for (int hw = 0; hw < 2; ++hw) {
            bool used[16] = {false};
            for (int l = 16 * hw; l < 16 * hw + 16; ++l) {
                const int m = 32 * sl + l;
                if (m >= c.n_mels) continue;
                const int b0 = bands[3 * m];
                int d = 0;
                while (d < 16 && d <= b0 && used[(b0 - d) & 15]) ++d;
                if (d == 16 || d > b0) d = 0;   // no free residue within reach: keep the band start (costs a wavefront)
                used[(b0 - d) & 15] = true;
                lead[l] = d;
                starts[m] = b0 - d;
            }
        }
        int km = 0;
        for (int l = 0; l < 32; ++l)
            if (32 * sl + l < c.n_mels) km = std::max(km, bands[3 * (32 * sl + l) + 1] + lead[l]);
        h->g_kmax[sl] = km;
        for (int i = 0; i < km; ++i)
            for (int l = 0; l < 32; ++l) {
                const int m = 32 * sl + l, j = i - lead[l];
                wts_t.push_back(m < c.n_mels && j >= 0 && j < bands[3 * m + 1] ? wts[(size_t)bands[3 * m + 2] + j] : 0.f);
            }
    }
    MS_CUDA(cudaMalloc(&h->d_gstarts_t, sizeof(int) * starts.size()));
    MS_CUDA(cudaMemcpy(h->d_gstarts_t, starts.data(), sizeof(int) * starts.size(), cudaMemcpyHostToDevice));
    h->n_gweights_t = (int)wts_t.size();
    if (wts_t.empty()) wts_t.push_back(0.f);
    MS_CUDA(cudaMalloc(&h->d_gweights_t, sizeof(float) * wts_t.size()));
    MS_CUDA(cudaMemcpy(h->d_gweights_t, wts_t.data(), sizeof(float) * wts_t.size(), cudaMemcpyHostToDevice));
    MS_CUDA(cudaMalloc(&h->d_gtw, sizeof(float2) * tw.size()));
    MS_CUDA(cudaMalloc(&h->d_gwin, sizeof(float) * std::max<size_t>(win.size(), 1)));
    MS_CUDA(cudaMalloc(&h->d_gbands, sizeof(int) * bands.size()));
    MS_CUDA(cudaMalloc(&h->d_gweights, sizeof(float) * wts.size()));
    MS_CUDA(cudaMemcpy(h->d_gtw, tw.data(), sizeof(float2) * tw.size(), cudaMemcpyHostToDevice));
    if (!win.empty()) MS_CUDA(cudaMemcpy(h->d_gwin, win.data(), sizeof(float) * win.size(), cudaMemcpyHostToDevice));
    MS_CUDA(cudaMemcpy(h->d_gbands, bands.data(), sizeof(int) * bands.size(), cudaMemcpyHostToDevice));
    MS_CUDA(cudaMemcpy(h->d_gweights, wts.data(), sizeof(float) * wts.size(), cudaMemcpyHostToDevice));
    h->mpl = (c.n_mels + 31) / 32;
    return MELSPEC_OK;
}

int32_t build_tables(melspec_handle* h) {
    if (h->plan == 1) return build_tables_generic(h);
    const Resolved& c = h->cfg;
    using namespace melspec;
    const int N = h->plan, nb = N / 2 + 1;
    const int R = N == 400 ? 20 : 32, C = N == 400 ? 20 : 16;
    std::vector<float> win;      // plan 400: [n1][t] float2 pairs; plan 512: [n1][c] floats
    std::vector<float4> tw;      // [i][t] = (W_N^(t*2i), W_N^(t*(2i+1)))
    std::vector<float2> rot(C);  // W_{2C}^(-c): pre-rotation of the middle row
    auto window_at = [&](int i) -> float {
        if (c.frontend == MELSPEC_FRONTEND_WHISPER)   // periodic Hann, reference src/stft.rs:141-145
            return (float)(0.5 * (1.0 - std::cos(2.0 * M_PI * (double)i / (double)N)));
        if (i >= c.frame_len) return 0.f;
        if (c.frontend == MELSPEC_FRONTEND_NEMO)      // symmetric Hann of win_length (src/mel.rs:708-719); its position inside
            return (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * (double)i / (double)(c.frame_len - 1)));   // n_fft only shifts phase
        // Povey window, zero padded to the FFT size (src/fbank.rs:100-105,184-190)
        return (float)std::pow(0.5 - 0.5 * std::cos(2.0 * M_PI * (double)i / (double)(c.frame_len - 1)), 0.85);
    };
    if (N == 400) {   // the kernel evaluates the periodic Hann window from these per-worker phase factors
        win.resize(40);
        for (int t = 0; t < 10; ++t) {
            const double th0 = 2.0 * M_PI * (double)(2 * t) / 400.0, th1 = 2.0 * M_PI * (double)(2 * t + 1) / 400.0;
            win[4 * t] = (float)std::cos(th0); win[4 * t + 1] = (float)std::cos(th1);
            win[4 * t + 2] = (float)std::sin(th0); win[4 * t + 3] = (float)std::sin(th1);
        }
    } else {
        win.resize(512);
        for (int i = 0; i < 512; ++i) win[i] = window_at(i);   // [n1][c] with i = 16*n1 + c
    }
    const int workers = R / 2, pairs = C / 2;
    tw.resize((size_t)pairs * workers);
    for (int t = 0; t < workers; ++t)
        for (int i = 0; i < pairs; ++i) {
            const double a0 = -2.0 * M_PI * (double)((t * 2 * i) % N) / (double)N;
            const double a1 = -2.0 * M_PI * (double)((t * (2 * i + 1)) % N) / (double)N;
            tw[(size_t)i * workers + t] = make_float4((float)std::cos(a0), (float)std::sin(a0), (float)std::cos(a1), (float)std::sin(a1));
        }
    for (int c2 = 0; c2 < C; ++c2) {
        const double a = 2.0 * M_PI * (double)c2 / (double)(2 * C);
        rot[c2] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    if (N == 400) {   // plan 400 reads one float4 per worker: (re c=2t, re c=2t+1, im c=2t, im c=2t+1)
        std::vector<float2> r4(20);
        for (int t = 0; t < 10; ++t) {
            r4[2 * t] = make_float2(rot[2 * t].x, rot[2 * t + 1].x);
            r4[2 * t + 1] = make_float2(rot[2 * t].y, rot[2 * t + 1].y);
        }
        rot = r4;
    }
    // sparse banded filterbank -> per-lane projection program
    for (int m = 0; m < c.n_mels; ++m)
        if (h->dense[(size_t)m * nb] != 0.0)
            return fail(MELSPEC_ERR_UNSUPPORTED, "filterbank has a non-zero DC column; the kernel never forms bin 0");
    struct Band { int mel; std::vector<std::pair<int, double>> e; };
    std::vector<Band> bands(c.n_mels);
    for (int m = 0; m < c.n_mels; ++m) {
        bands[m].mel = m;
        // Whisper drops bins >= N/2 (reference src/mel.rs:158-162)
        const int last = c.frontend == MELSPEC_FRONTEND_WHISPER ? N / 2 - 1 : N / 2;
        for (int b = 1; b <= last; ++b) {
            const double w = h->dense[(size_t)m * nb + b];
            if (w != 0.0) bands[m].e.push_back({b, w});
        }
    }
    std::vector<float2> proj;
    std::vector<int> meta(kMetaInts, -1);
    h->mpl = (c.n_mels + 31) / 32;
    {
        // ---- windowed projection program (both plans).  A lane's slot-s entries are K_s consecutive bins [start, start + K_s)
        // of the bin-ordered power rows; the table holds the weights only.  Plan 400 reads 8-byte rows of three planes with
        // LDS.64 (conflicts are per half-warp, rows mod 16), plan 512 reads 16-byte rows with LDS.128 (per quarter-warp, mod 8).
        const int grp = N == 400 ? 16 : 8, ngrp = 32 / grp, last_row = N / 2;
        // A piece is a run of consecutive bins of one mel band.  Long bands may be split into two pieces that sit in the same
        // slot on two lanes and are added up with one warp shuffle per frame (plan 400): the slot lengths K_s are set by the
        // longest piece, so halving the long bands shortens every lane's loop (Whisper-80: 14+4+2 -> 8+5+2 entries).
        struct Piece { int mel, b0, len, emit, pair; };   // mel: weight row; emit: this lane stores the mel; pair: id or -1
        // Measured on B200 (profiles/README.md): 23 fewer shared-memory wavefronts per pass, but the 12 shuffles and their
        // latency cost more than that (0.4245 -> 0.4330 ms), so splitting is opt-in (MELSPEC_SPLIT=1) and off by default.
        static const bool split_enabled = [] { const char* e = std::getenv("MELSPEC_SPLIT"); return e && e[0] == '1'; }();
        const int capacity = 32 * h->mpl;
        std::vector<Piece> pcs;
        int K[kMaxMpl] = {0, 0, 0, 0}, EX[kMaxMpl] = {0, 0, 0, 0}, ktot = 0;
        auto build = [&](int L, std::vector<Piece>& out, int (&Kc)[kMaxMpl], int (&Ex)[kMaxMpl]) -> int {
            struct Unit { Piece a, b; int n, key; };
            std::vector<Unit> units;
            int npieces = 0, pid = 0;
            for (int m = 0; m < c.n_mels; ++m) {
                Piece pc{m, 1, 0, 1, -1};
                if (!bands[m].e.empty()) { pc.b0 = bands[m].e.front().first; pc.len = bands[m].e.back().first - pc.b0 + 1; }
                if (L > 0 && pc.len > L) {
                    const int hlen = (pc.len + 1) / 2;
                    Piece a{m, pc.b0, hlen, 1, pid}, b2{m, pc.b0 + hlen, pc.len - hlen, 0, pid};
                    ++pid;
                    units.push_back(Unit{a, b2, 2, hlen});
                    npieces += 2;
                } else {
                    units.push_back(Unit{pc, pc, 1, pc.len});
                    npieces += 1;
                }
            }
            if (npieces > capacity) return -1;
            std::stable_sort(units.begin(), units.end(), [](const Unit& x, const Unit& y) { return x.key > y.key; });
            std::vector<std::vector<Piece>> slots(h->mpl);
            for (const Unit& u : units) {
                bool placed = false;
                for (int sl = 0; sl < h->mpl && !placed; ++sl)
                    if ((int)slots[sl].size() + u.n <= 32) {
                        slots[sl].push_back(u.a);
                        if (u.n == 2) slots[sl].push_back(u.b);
                        placed = true;
                    }
                if (!placed) return -1;
            }
            out.clear();
            int tot = 0;
            for (int sl = 0; sl < kMaxMpl; ++sl) { Kc[sl] = 0; Ex[sl] = 0; }
            for (int sl = 0; sl < h->mpl; ++sl) {
                for (const Piece& pc : slots[sl]) { Kc[sl] = std::max(Kc[sl], pc.len); Ex[sl] |= pc.pair >= 0; }
                while ((int)slots[sl].size() < 32) slots[sl].push_back(Piece{-1, 1, 0, 0, -1});   // idle lanes
                out.insert(out.end(), slots[sl].begin(), slots[sl].end());
                tot += Kc[sl];
            }
            return tot;
        };
        {
            int best = build(0, pcs, K, EX), bestL = 0;
            if (N == 400 && split_enabled) {
                int maxlen = 0;
                for (int m = 0; m < c.n_mels; ++m)
                    if (!bands[m].e.empty()) maxlen = std::max(maxlen, bands[m].e.back().first - bands[m].e.front().first + 1);
                for (int L = maxlen - 1; L >= 3; --L) {
                    std::vector<Piece> t;
                    int Kt[kMaxMpl], Et[kMaxMpl];
                    const int tot = build(L, t, Kt, Et);
                    if (tot < 0) continue;
                    int nex = 0;
                    for (int sl = 0; sl < kMaxMpl; ++sl) nex += Et[sl];
                    if (tot + nex < best) { best = tot + nex; bestL = L; }   // an exchange costs about one entry
                }
                if (bestL) build(bestL, pcs, K, EX);
            }
            ktot = 0;
            for (int sl = 0; sl < h->mpl; ++sl) { ktot += K[sl]; meta[sl] = K[sl] | (EX[sl] << 16); }
            for (int sl = h->mpl; sl < kMaxMpl; ++sl) meta[sl] = 0;
        }
        // (1) The output rows are staged with one 32-bit store per (slot, frame): conflict-free when the 32 mels of a slot
        // differ mod 32.  Unsplit pieces may change slots as long as they still fit (len <= K of the new slot).
        auto stage_cost = [&]() {
            int tot = 0;
            for (int s = 0; s < h->mpl; ++s) {
                int cnt[32] = {0};
                for (int l = 0; l < 32; ++l) {
                    const Piece& pc = pcs[(size_t)32 * s + l];
                    if (pc.mel >= 0 && pc.emit) tot += cnt[pc.mel & 31]++;
                }
            }
            return tot;
        };
        uint64_t rng = 0x9E3779B97F4A7C15ull;
        auto rnd = [&rng](int n) {
            rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
            return (int)((rng >> 33) % (uint64_t)n);
        };
        int sc = stage_cost();
        for (int iter = 0; iter < 20000 && sc > 0; ++iter) {
            const int a = rnd(32 * h->mpl), b = rnd(32 * h->mpl);
            if (a / 32 == b / 32 || pcs[a].len > K[b / 32] || pcs[b].len > K[a / 32] || pcs[a].pair >= 0 || pcs[b].pair >= 0) continue;
            std::swap(pcs[a], pcs[b]);
            const int nc = stage_cost();
            if (nc <= sc) sc = nc; else std::swap(pcs[a], pcs[b]);
        }
        // (2) Window starts: piece i may start anywhere in [max(1, b0 + len - K), min(b0, N/2 + 1 - K)] (rows 1..N/2 are the ones
        // the kernel writes every pass).  The power loads of an entry are conflict-free when the `grp` lanes of a lane group
        // start at `grp` different rows mod `grp`: a perfect matching of the slot's 32 pieces onto (lane group, residue) pairs
        // (Kuhn's augmenting paths; whatever stays unmatched is placed anyway and merely costs a wavefront).
        const int ktot4 = std::max(4, (ktot + 3) / 4 * 4);
        std::vector<float> wtab((size_t)2 * ktot4 * 32, 0.f);
        int eoff = 0;
        h->proj_wavefront_cost = 0;
        for (int s = 0; s < h->mpl; ++s) {
            const int Ks = K[s];
            int lo[32], hi[32];
            for (int i = 0; i < 32; ++i) {
                const Piece& pc = pcs[(size_t)32 * s + i];
                lo[i] = std::max(1, pc.b0 + pc.len - Ks);
                hi[i] = std::min(pc.b0, last_row + 1 - Ks);
                if (pc.len == 0) { lo[i] = 1; hi[i] = std::max(1, last_row + 1 - Ks); }
                if (hi[i] < lo[i]) hi[i] = lo[i];
            }
            int owner[32];   // (half-warp, residue) -> piece
            std::fill(owner, owner + 32, -1);
            std::function<bool(int, std::vector<char>&)> augment = [&](int i, std::vector<char>& seen) {
                for (int st = lo[i]; st <= hi[i] && st < lo[i] + grp; ++st)
                    for (int hw = 0; hw < ngrp; ++hw) {
                        const int node = grp * hw + (st & (grp - 1));
                        if (seen[node]) continue;
                        seen[node] = 1;
                        if (owner[node] < 0 || augment(owner[node], seen)) { owner[node] = i; return true; }
                    }
                return false;
            };
            std::vector<int> order(32);
            for (int i = 0; i < 32; ++i) order[i] = i;
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return hi[a] - lo[a] < hi[b] - lo[b]; });
            std::vector<int> unmatched;
            for (int i : order) {
                std::vector<char> seen(32, 0);
                if (!augment(i, seen)) unmatched.push_back(i);
            }
            int lane_piece[32], lane_start[32];
            std::fill(lane_piece, lane_piece + 32, -1);
            int fill[4] = {0, 0, 0, 0};
            for (int node = 0; node < 32; ++node) {
                const int i = owner[node];
                if (i < 0) continue;
                const int hw = node / grp, l = grp * hw + fill[hw]++;
                int st = lo[i];
                while ((st & (grp - 1)) != (node & (grp - 1))) ++st;
                lane_piece[l] = i; lane_start[l] = st;
            }
            for (int i : unmatched) {
                int hw = 0;
                while (hw < ngrp - 1 && fill[hw] >= grp) ++hw;
                const int l = grp * hw + fill[hw]++;
                lane_piece[l] = i; lane_start[l] = lo[i];
                h->proj_wavefront_cost += 3 * Ks;
            }
            for (int l = 0; l < 32; ++l) {   // partner lane of a split band's piece (own lane otherwise)
                const Piece& pc = pcs[(size_t)32 * s + lane_piece[l]];
                int partner = l;
                if (pc.pair >= 0)
                    for (int l2 = 0; l2 < 32; ++l2)
                        if (l2 != l && pcs[(size_t)32 * s + lane_piece[l2]].pair == pc.pair) partner = l2;
                meta[kMaxMpl + 2 * kMaxMpl * 32 + s * 32 + l] = partner;
            }
            for (int l = 0; l < 32; ++l) {
                const Piece& pc = pcs[(size_t)32 * s + lane_piece[l]];
                meta[kMaxMpl + s * 32 + l] = pc.emit ? pc.mel : -1;
                const int plane_unit = lane_start[l];
                meta[kMaxMpl + kMaxMpl * 32 + s * 32 + l] = plane_unit;
                for (int e = 0; e < Ks; ++e) {
                    const int bin = lane_start[l] + e;
                    const float w = (pc.mel >= 0 && bin < nb) ? (float)(h->dense[(size_t)pc.mel * nb + bin] * 0.25) : 0.f;
                    const bool inband = pc.len > 0 && bin >= pc.b0 && bin < pc.b0 + pc.len;
                    const int ge = eoff + e;
                    wtab[(size_t)ge * 32 + l] = inband ? w : 0.f;
                    wtab[(size_t)ktot4 * 32 + ((size_t)(ge >> 2) * 32 + l) * 4 + (ge & 3)] = inband ? w : 0.f;
                }
            }
            h->proj_wavefront_cost += 2 * 3 * Ks;
            eoff += Ks;
        }
        for (int s = h->mpl; s < kMaxMpl; ++s)
            for (int l = 0; l < 32; ++l) { meta[kMaxMpl + kMaxMpl * 32 + s * 32 + l] = 1; meta[kMaxMpl + 2 * kMaxMpl * 32 + s * 32 + l] = l; }
        h->proj_ktot = ktot4;
        h->kspec = 0;
        if (N == 400 && h->mpl == 3) {
            if (K[0] == 14 && K[1] == 4 && K[2] == 2 && !EX[0] && !EX[1] && !EX[2]) h->kspec = 1;
            if (K[0] == 8 && K[1] == 5 && K[2] == 2 && EX[0] && EX[1] && !EX[2]) h->kspec = 2;
        }
        if (N == 400 && h->mpl == 4 && K[0] == 9 && K[1] == 4 && K[2] == 2 && K[3] == 1 && !EX[0] && !EX[1] && !EX[2] && !EX[3]) h->kspec = 4;
        h->ksched512 = 0;
        if (N == 512 && !EX[0] && !EX[1] && !EX[2] && !EX[3])
            for (int k = 1; k < 5; ++k)
                if (K[0] == melspec::ksched512(k, 0) && K[1] == melspec::ksched512(k, 1) && K[2] == melspec::ksched512(k, 2) &&
                    K[3] == melspec::ksched512(k, 3))
                    h->ksched512 = k;
        proj.resize(wtab.size() / 2);
        std::memcpy(proj.data(), wtab.data(), wtab.size() * sizeof(float));
    }
    MS_CUDA(cudaMalloc(&h->d_window, sizeof(float) * win.size()));
    MS_CUDA(cudaMalloc(&h->d_twiddle, sizeof(float4) * tw.size()));
    MS_CUDA(cudaMalloc(&h->d_rot10, sizeof(float2) * rot.size()));
    MS_CUDA(cudaMalloc(&h->d_proj, sizeof(float2) * proj.size()));
    MS_CUDA(cudaMalloc(&h->d_meta, sizeof(int) * meta.size()));
    MS_CUDA(cudaMemcpy(h->d_window, win.data(), sizeof(float) * win.size(), cudaMemcpyHostToDevice));
    MS_CUDA(cudaMemcpy(h->d_twiddle, tw.data(), sizeof(float4) * tw.size(), cudaMemcpyHostToDevice));
    MS_CUDA(cudaMemcpy(h->d_rot10, rot.data(), sizeof(float2) * rot.size(), cudaMemcpyHostToDevice));
    MS_CUDA(cudaMemcpy(h->d_proj, proj.data(), sizeof(float2) * proj.size(), cudaMemcpyHostToDevice));
    MS_CUDA(cudaMemcpy(h->d_meta, meta.data(), sizeof(int) * meta.size(), cudaMemcpyHostToDevice));
    return MELSPEC_OK;
}

// Warps per CTA (one persistent CTA per SM).  Measured best: 12 for both plans (168 registers/thread, three warps per
// scheduler hide the fixed 2-cycle issue cadence of the packed FADD2/FFMA2 stream better than two).  MELSPEC_WARPS=8|12 overrides for tuning; the count must be a multiple of 4
// (registers are allocated per SM sub-partition).
int warps_per_cta(int plan) {
    static int forced = [] {
        const char* e = std::getenv("MELSPEC_WARPS");
        const int v = e ? std::atoi(e) : 0;
        return (v == 8 || v == 12) ? v : 0;   // (16: see launch_device, plan 400 only)
    }();
    (void)plan;
    return forced ? forced : 12;
}

template <typename Kern>
int32_t launch_kernel(Kern kern, const melspec::KParams& p, int grid, int threads, size_t smem, cudaStream_t st) {
    // the opt-in for > 48 KB of dynamic shared memory is per function and per device; it is cheap, so set it every time
    MS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    kern<<<grid, threads, smem, st>>>(p);
    MS_CUDA(cudaGetLastError());
    return MELSPEC_OK;
}

// Kernels that couple all frames of a clip and therefore follow the fused kernel: NeMo per-feature mean/std
// (src/mel.rs:721-749) and Kaldi CMN (src/fbank.rs:226-233) when it is not fused.
int32_t launch_post_kernels(melspec_handle* h, const melspec::KParams& p, int64_t n_clips, const int32_t* d_lens, float* d_out,
                            bool cmn_done, cudaStream_t st) {
    using namespace melspec;
    const Resolved& c = h->cfg;
    if (c.frontend == MELSPEC_FRONTEND_NEMO && c.norm_feat) {
        const long long rows = (long long)n_clips * c.n_mels;
        melspec_featnorm_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(d_out, p.out_clip_stride, p.out_row_stride,
                                                                          p.frames_per_clip, c.n_mels, (int)rows, d_lens, p.n_samples,
                                                                          c.hop, c.fft, c.center);
        MS_CUDA(cudaGetLastError());
        h->launches += 1;
    }
    if (c.frontend == MELSPEC_FRONTEND_KALDI && c.cmn && !cmn_done) {
        melspec_cmn_kernel<<<(unsigned)n_clips, 512, 0, st>>>(d_out, p.out_clip_stride, p.frames_per_clip, c.n_mels, d_lens,
                                                             p.n_samples, c.frame_len, c.hop);
        MS_CUDA(cudaGetLastError());
        h->launches += 1;
    }
    return MELSPEC_OK;
}

// The general plan, pair form (melspec_generic2.cuh): one warp per two neighbouring frames.  Returns MELSPEC_OK with *launched = false
// when two frames' buffers would leave too few warps on an SM; the caller then runs the one-frame kernel.
template <int NFT, bool INPLACE = false, bool LB = false>
int32_t launch_generic_pair_t(melspec_handle* h, const melspec::KParams& p, const melspec::GParams& g, int64_t n_clips, cudaStream_t st,
                              bool* launched) {
    using namespace melspec;
    auto kern = melspec_generic_pair_kernel<NFT, INPLACE, LB>;
    *launched = false;
    if (g.n_weights_t == 0) return MELSPEC_OK;
    const size_t w_bytes = ((size_t)4 * g.n_weights_t + 15) & ~(size_t)15;
    const size_t stw_bytes = ((size_t)8 * g2_stw_elems(NFT ? NFT : 1) + 15) & ~(size_t)15;
    const size_t budget = 226 * 1024, per_warp = (size_t)(INPLACE ? 16 : 32) * generic2_buf_elems(g.Nf),
                 tw_bytes = (((size_t)8 * (NFT ? g.N / 4 + 1 : g.N) + 15) & ~(size_t)15) + w_bytes + stw_bytes;
    // warps per CTA: the count that keeps the most warps resident on an SM (one twiddle / weight table per CTA), asked of the
    // occupancy calculator itself (register allocation granularity decides between one and two CTAs); 8 unless another count is
    // clearly better.  Computed once per handle.
    MS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (h->pair_nw == 0) {
        int warps_at[17] = {0};
        for (int cand = 2; cand <= (LB ? 8 : 16); ++cand) {
            const size_t need = tw_bytes + per_warp * cand;
            if (need > budget) break;
            int blocks = 0;
            MS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kern, cand * 32, need));
            warps_at[cand] = cand * blocks;
        }
        int nw = 8;
        while (nw > 2 && warps_at[nw] == 0) --nw;
        for (int cand = 2; cand <= 16; ++cand)
            if (warps_at[cand] * 100 > warps_at[nw] * (cand < nw ? 125 : 110)) nw = cand;
        h->pair_nw = warps_at[nw] > 0 ? nw : -1;
        h->pair_warps = warps_at[nw];
    }
    static const int min_warps = [] { const char* e = std::getenv("MELSPEC_PAIR_MIN_WARPS"); return e ? std::atoi(e) : 8; }();
    if (h->pair_nw < 0 || h->pair_warps < min_warps) return MELSPEC_OK;   // large transforms: the one-frame kernel keeps more warps in flight
    const int nw = h->pair_nw, per_sm = h->pair_warps / nw;
    const size_t smem = tw_bytes + per_warp * nw;
    const long long n_pairs = (long long)((p.frames_per_clip + 1) / 2) * n_clips;
    const long long want = (n_pairs + nw - 1) / nw;
    const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)h->num_sms * per_sm));
    kern<<<grid, nw * 32, smem, st>>>(p, g);
    MS_CUDA(cudaGetLastError());
    *launched = true;
    return MELSPEC_OK;
}

int32_t launch_generic_pair(melspec_handle* h, const melspec::KParams& p, const melspec::GParams& g, int64_t n_clips, cudaStream_t st,
                            bool* launched) {
    // MELSPEC_GENERIC_PAIR: 0 = one-frame kernel only, 1 (default) = pair form, compiled-in sizes where they exist, 2 = pair form with
    // sizes read at run time only, 3 = pair form for the compiled-in sizes only
    static const int mode = [] { const char* e = std::getenv("MELSPEC_GENERIC_PAIR"); return e ? std::atoi(e) : 1; }();
    // MELSPEC_PAIR_INPLACE=0: two buffers per warp everywhere (A/B); default: one buffer where two leave fewer warps than the registers allow
    static const bool inplace = [] { const char* e = std::getenv("MELSPEC_PAIR_INPLACE"); return !(e && e[0] == '0'); }();
    static const bool lb = [] { const char* e = std::getenv("MELSPEC_PAIR_LB"); return !(e && e[0] == '0'); }();   // (0: A/B against the 128-register build)
    *launched = false;
    if (mode == 0) return MELSPEC_OK;
    if (mode == 1 && g.N == 2 * g.Nf) {
        switch (g.Nf) {
            // measured (profiles/r2_generic_pair_ab.txt): the 80-register build pays at fft 256 only (945 vs 894 M frames/s; fft 480 /
            // 512: 5 - 7 % slower); the one-buffer form pays at fft 1024 (476 vs 407 M) and loses at fft 800 (spills: 309 vs 421 M)
            case 128: return lb ? launch_generic_pair_t<128, false, true>(h, p, g, n_clips, st, launched)
                                : launch_generic_pair_t<128>(h, p, g, n_clips, st, launched);
            case 240: return launch_generic_pair_t<240>(h, p, g, n_clips, st, launched);
            case 256: return launch_generic_pair_t<256>(h, p, g, n_clips, st, launched);
            case 320: return launch_generic_pair_t<320>(h, p, g, n_clips, st, launched);
            case 400: return launch_generic_pair_t<400>(h, p, g, n_clips, st, launched);
            case 512: return inplace ? launch_generic_pair_t<512, true>(h, p, g, n_clips, st, launched)
                                     : launch_generic_pair_t<512>(h, p, g, n_clips, st, launched);
            default: break;
        }
    }
    // other sizes: the pair form with sizes read at run time (measured 1.1 - 1.4 x the one-frame kernel; the compiled-in sizes 1.4 - 1.9 x)
    if (mode == 3) return MELSPEC_OK;   // (A/B: compiled-in sizes only)
    return launch_generic_pair_t<0>(h, p, g, n_clips, st, launched);
}

// The general plan: one warp per frame (or per two frames, above), mixed-radix shared-memory FFT (melspec_generic.cuh).
int32_t launch_generic(melspec_handle* h, melspec::KParams& p, int64_t n_clips, const int32_t* d_lens, float* d_out,
                       int64_t row_stride, cudaStream_t st) {
    using namespace melspec;
    const Resolved& c = h->cfg;
    GParams g{};
    g.tw = h->d_gtw; g.window = h->d_gwin; g.bands = h->d_gbands; g.weights = h->d_gweights;
    g.N = c.fft;
    g.Nf = (c.fft % 2 == 0) ? c.fft / 2 : c.fft;
    g.n_stages = (int)h->radices.size();
    for (int i = 0, st = 1; i < g.n_stages; ++i) {
        g.radix[i] = h->radices[i];
        int sh = -1;
        if ((st & (st - 1)) == 0) { sh = 0; while ((1 << sh) < st) ++sh; }
        g.sshift[i] = sh;
        st *= h->radices[i];
    }
    g.mode = c.frontend == MELSPEC_FRONTEND_KALDI ? 1 : c.frontend == MELSPEC_FRONTEND_NEMO ? 2 : 0;
    g.use_power = c.use_power; g.use_log = c.use_log; g.center = c.center;
    g.n_units = (long long)p.frames_per_clip * n_clips;
    g.weights_t = h->d_gweights_t; g.starts_t = h->d_gstarts_t;
    g.n_weights_t = h->n_gweights_t * 4 <= 32 * 1024 ? h->n_gweights_t : 0;   // (0: too large for shared memory, no pair form)
    for (int sl = 0; sl < 4; ++sl) g.kmax[sl] = h->g_kmax[sl];
    p.n_clips = (int)n_clips;
    // warps per CTA: as many as fit beside the twiddle table (8 N bytes) at 16 Nf bytes each, at most 8
    g.vec2 = ((uintptr_t)p.pcm % 8 == 0) && (p.clip_stride % 2 == 0) && (c.hop % 2 == 0) && (p.frame_offset % 2 == 0) && (c.fft % 2 == 0);
    if (c.frontend == MELSPEC_FRONTEND_NEMO && row_stride > p.frames_per_clip && p.out_clip_stride == row_stride * c.n_mels) {
        const long long total = (long long)n_clips * c.n_mels * (row_stride - p.frames_per_clip);   // pad_to columns are zeros (src/mel.rs:336)
        melspec_zero_cols_kernel<<<(int)std::min<long long>((total + 255) / 256, (long long)h->num_sms * 8), 256, 0, st>>>(
            d_out, p.out_clip_stride, c.n_mels, row_stride, p.frames_per_clip, (int)row_stride, n_clips);
        MS_CUDA(cudaGetLastError());
    }
    {
        bool launched = false;
        const int32_t rc = launch_generic_pair(h, p, g, n_clips, st, &launched);
        if (rc != MELSPEC_OK) return rc;
        if (launched) {
            h->launches += 1;
            return launch_post_kernels(h, p, n_clips, d_lens, d_out, false, st);
        }
    }
    const size_t budget = 220 * 1024, per_warp = (size_t)16 * generic_buf_elems(g.Nf),
                 tw_bytes = (size_t)8 * c.fft + (size_t)4 * ((g.n_weights_t + 1) & ~1);   // twiddles + the transposed band weights
    // warps per CTA: the count that puts the most warps on an SM (the kernel is latency bound: every FFT stage is a
    // round trip through shared memory), given one twiddle table per CTA, 227 KB of shared memory and 64 K registers per SM
    cudaFuncAttributes fa;
    MS_CUDA(cudaFuncGetAttributes(&fa, melspec_generic_kernel));
    const int reg_warps = std::max(1, 65536 / (std::max(fa.numRegs, 1) * 32));
    int warps_at[17] = {0};
    for (int cand = 1; cand <= 16; ++cand) {
        const size_t need = tw_bytes + per_warp * cand + 1024;   // + the per-CTA reservation
        if (need > budget + 1024) break;
        const int ctas = (int)std::min<size_t>(32, (228 * 1024) / need);
        warps_at[cand] = std::min(std::min(cand * ctas, 64), reg_warps / cand * cand);
    }
    // 8-warp CTAs by default (neighbouring frames of a CTA share their PCM in L1; measured: smaller CTAs lose more than their
    // extra warps gain); another size only where it puts clearly more warps on the SM (large transforms)
    int nw = 8;
    while (nw > 1 && warps_at[nw] == 0) --nw;
    for (int cand = 1; cand <= 16; ++cand)
        if (warps_at[cand] * 100 > warps_at[nw] * (cand < nw ? 140 : 115)) nw = cand;
    const size_t smem = tw_bytes + per_warp * nw;
    MS_CUDA(cudaFuncSetAttribute(melspec_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    int per_sm = 1;
    MS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, melspec_generic_kernel, nw * 32, smem));
    if (per_sm < 1) return fail(MELSPEC_ERR_UNSUPPORTED, "fft_size too large for the shared-memory FFT of this build");
    const long long want = (g.n_units + nw - 1) / nw;
    const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)h->num_sms * per_sm));
    melspec_generic_kernel<<<grid, nw * 32, smem, st>>>(p, g);
    MS_CUDA(cudaGetLastError());
    h->launches += 1;
    return launch_post_kernels(h, p, n_clips, d_lens, d_out, false, st);
}

// Core launch: device pointers, explicit frame count (frames_per_clip may be smaller than num_frames(n_samples)).
int32_t launch_device(melspec_handle* h, const float* d_pcm, int64_t n_clips, int64_t clip_stride, int64_t n_samples,
                      int64_t frames_per_clip, const int32_t* d_lens, float* d_out, int64_t out_clip_stride, int32_t layout,
                      cudaStream_t st, int64_t row_stride_override = 0) {
    using namespace melspec;
    const Resolved& c = h->cfg;
    if (n_clips == 0 || frames_per_clip == 0) return MELSPEC_OK;
    if (n_samples > 0x7fffffff) return fail(MELSPEC_ERR_INVALID_ARG, "n_samples per clip must fit in int32");
    const bool kaldi = c.frontend == MELSPEC_FRONTEND_KALDI, nemo = c.frontend == MELSPEC_FRONTEND_NEMO;
    if (kaldi && layout != MELSPEC_LAYOUT_FRAME_MAJOR)
        return fail(MELSPEC_ERR_UNSUPPORTED, "the Kaldi frontend produces (T, n_mels) frame-major output only");
    if (nemo && layout != MELSPEC_LAYOUT_MEL_MAJOR)
        return fail(MELSPEC_ERR_UNSUPPORTED, "the NeMo frontend produces (n_mels, frames) feature-major output only");
    const int64_t row_stride = row_stride_override > 0 ? row_stride_override : nemo ? padded_frames_for(c, n_samples) : frames_per_clip;
    const int fpw = h->plan == 400 ? p400::FPW : h->plan == 512 ? p512::FPW : 2;
    KParams p{};
    p.pcm = d_pcm; p.out = d_out; p.lens = d_lens;
    p.clip_stride = clip_stride;
    p.out_clip_stride = out_clip_stride ? out_clip_stride : row_stride * c.n_mels;
    p.out_row_stride = (int)row_stride;
    p.frame_offset = nemo ? (c.fft - c.frame_len) / 2 - (c.center ? c.fft / 2 : 0) : 0;
    p.log_add = nemo ? (float)c.guard : 0.f;
    p.n_samples = (int)n_samples;
    p.frames_per_clip = (int)frames_per_clip;
    p.wtiles_per_clip = (int)((frames_per_clip + fpw - 1) / fpw);
    const int64_t n_wtiles = (int64_t)p.wtiles_per_clip * n_clips;
    if (n_wtiles > 0x7fffffff - 148 * 64) return fail(MELSPEC_ERR_INVALID_ARG, "too many frames for one launch");
    p.n_wtiles = (int)n_wtiles;
    p.hop = c.hop; p.n_mels = c.n_mels; p.fft_size = c.fft; p.layout = layout;
    p.frame_len = c.frame_len; p.preemph = (float)c.preemph;
    const bool hop160 = c.hop == 160;
    const bool aligned_in = ((uintptr_t)d_pcm % 16 == 0) && (clip_stride % 4 == 0) && (n_samples % 4 == 0) && (c.hop % 4 == 0);
    p.bulk_in = aligned_in ? 1 : 0;
    p.bulk_out = (layout == MELSPEC_LAYOUT_FRAME_MAJOR) && ((uintptr_t)d_out % 16 == 0) && (c.n_mels % 4 == 0) &&
                 (p.out_clip_stride % 4 == 0);
    p.mm_aligned8 = (layout == MELSPEC_LAYOUT_MEL_MAJOR) && ((uintptr_t)d_out % 8 == 0) && (p.out_clip_stride % 2 == 0) && (row_stride % 2 == 0);
    p.window = reinterpret_cast<const float2*>(h->d_window); p.twiddle = h->d_twiddle; p.rot10 = h->d_rot10;
    p.proj = h->d_proj; p.proj_meta = h->d_meta; p.proj_ktot = h->proj_ktot;
    p.floor_val = (float)c.floor;
    if (p.floor_val > 0.f && p.floor_val < 1.17549435e-38f) p.floor_val = 1.17549435e-38f;   // lg2_normal() flushes denormals
    {   // pair prescale (melspec_kernels.cuh): a frame may be scaled by 2^k as long as floor * 2^(2k) (NeMo: guard * 2^(2k)) stays normal
        const float v = nemo ? p.log_add : p.floor_val;
        uint32_t bits;
        std::memcpy(&bits, &v, 4);
        const int ef = (int)((bits >> 23) & 255u);
        p.ps_down = v > 0.f ? std::min(melspec::kMaxShift, std::max(0, (ef - 1) / 2)) : 0;
        p.ps_up = v > 0.f ? std::min(melspec::kMaxShift, std::max(0, (254 - ef) / 2)) : 0;
        static const bool no_prescale = [] { const char* e = std::getenv("MELSPEC_NO_PRESCALE"); return e && e[0] == '1'; }();   // (debugging / A-B only)
        if (no_prescale) p.ps_down = p.ps_up = 0;
    }
    if (kaldi) { p.log_mul = c.use_log ? (float)std::log(2.0) : 0.f; p.normalize = 0; }   // ln(max(e, floor)), src/fbank.rs:207-221
    else if (nemo) { p.log_mul = (float)std::log(2.0); p.normalize = 0; }                 // ln(e + guard), src/mel.rs:365-368
    else { p.log_mul = (float)std::log10(2.0); p.normalize = 1; }                         // log10 + per-frame clamp, src/mel.rs:148-168,645-654
    if (h->plan == 1) return launch_generic(h, p, n_clips, d_lens, d_out, row_stride, st);
    // shared-memory carve-up: [mbarriers | window | twiddles | projection program | meta | per-warp slabs]
    auto up = [](size_t v, size_t a) { return (v + a - 1) / a * a; };
    size_t off = 128;
    p.smem_win = (int)off; off = up(off + (h->plan == 400 ? 1600 : 2048), 128);
    p.smem_tw = (int)off; off = up(off + (h->plan == 400 ? 4800 : 2048), 128);   // plan 400: one copy per FFT of the warp
    p.smem_rot = (int)off; off = up(off + 512, 128);
    p.smem_proj = (int)off; off = up(off + sizeof(float2) * 32 * (size_t)(h->proj_ktot + 1), 128);
    p.smem_meta = (int)off; off = up(off + sizeof(int) * kMetaInts, 128);
    p.smem_mmoff = (int)off; if (h->plan == 400) off = up(off + sizeof(int) * 3 * 128, 128);
    p.smem_cmn = (int)off;
    // fused CMN: mode 1 column sums per warp + means; mode 2 two sets of column sums + kCmnRows replicated rows of -mean
    if (h->plan == 512) off = up(off + sizeof(float) * 128 * (size_t)(kaldi && c.cmn ? 2 * 12 + p512::kCmnRows : 12 + 1), 128);
    p.smem_warp0 = (int)off;
    static_assert(p400::FPW * 32 * kMaxMpl * 4 <= p400::STAGE_MAX, "output rows must fit behind the power rows");
    static_assert(p512::FPW * 32 * kMaxMpl * 4 <= p512::STAGE_MAX, "output rows must fit behind the power rows");
    size_t pcm_words;
    int nw = warps_per_cta(h->plan);
    if (h->plan == 400) {
        pcm_words = hop160 ? (size_t)p400::NCHUNK * p400::CS320 : (size_t)(p400::FPW - 1) * c.hop + 400;
        p.smem_stage_off = p400::PBYTES;
        // 16 warps per SM (four per scheduler) for the headline configuration (Whisper 80-mel, hop 160): 111 registers per thread
        // with the twiddles read from shared memory, and a per-warp footprint of 12.4 KB instead of 16 KB because the PCM stage
        // starts inside the exchange slab, right behind the output rows (the refill is issued once the slab's tail is dead).
        // Measured (profiles/README.md, r2): 0.4017 ms against 0.4010 ms with 12 warps -- the fourth warp per scheduler buys nothing
        // because the kernel is bound by the shared-memory pipe, not by latency -- so it is opt-in (MELSPEC_WARPS=16).
        static const bool w16 = [] { const char* e = std::getenv("MELSPEC_WARPS"); return e && std::atoi(e) == 16; }();
        if (w16 && hop160 && h->kspec == 1 && h->mpl == 3) nw = 16;
        if (nw == 16) {
            p.smem_pcm_off = p400::PBYTES + p400::FPW * 32 * 3 * 4;               // behind the staged output rows (80 mels: 6864)
            p.smem_scr_off = (int)up((size_t)p.smem_pcm_off + pcm_words * 4, 16);   // 48 bytes: the pair prescale's (floor, log offset)
            p.smem_warp_stride = (int)up((size_t)p.smem_scr_off + 48, 128);
        } else {
            p.smem_scr_off = p400::ZBYTES;
            p.smem_pcm_off = (int)up(p400::ZBYTES + 48, 128);
            p.smem_warp_stride = (int)up((size_t)p.smem_pcm_off + pcm_words * 4, 128);
        }
    } else {
        pcm_words = (size_t)p512::NCHUNK * p512::CS;
        p.smem_stage_off = p512::PBYTES;
        p.smem_scr_off = p512::ZBYTES;
        p.smem_pcm_off = (int)up(p512::ZBYTES + p512::SCRBYTES, 128);
        p.smem_warp_stride = (int)up((size_t)p.smem_pcm_off + pcm_words * 4, 128);
    }
    // Kaldi CMN inside the fused kernel: one CTA works through whole clips (see melspec512_kernel).  Needs enough clips to
    // fill the GPU, enough tiles per clip for every warp, and float4-addressable rows; otherwise CMN stays a second kernel.
    // MELSPEC_CMN_FUSED: 0 = CMN as a second kernel, 1 = block barrier + in-place subtraction by the CTA (round 1), 2 (default) =
    // no barrier, subtraction by TMA bulk reductions at the L2
    static const int cmn_fuse_mode = [] { const char* e = std::getenv("MELSPEC_CMN_FUSED"); return e && e[0] >= '0' && e[0] <= '2' ? e[0] - '0' : 2; }();
    const bool cmn_fuse_enabled = cmn_fuse_mode != 0;
    p.n_clips = (int)n_clips;
    p.vec_out = (layout == MELSPEC_LAYOUT_FRAME_MAJOR) && ((uintptr_t)d_out % 16 == 0) && (c.n_mels % 4 == 0) && (p.out_clip_stride % 4 == 0);
    const bool fused_cmn = cmn_fuse_enabled && kaldi && c.cmn && h->plan == 512 && layout == MELSPEC_LAYOUT_FRAME_MAJOR && p.vec_out &&
                           n_clips >= h->num_sms && p.wtiles_per_clip >= 2 * nw && c.n_mels <= 128;
    p.cmn_fused = fused_cmn ? cmn_fuse_mode : 0;
    if (fused_cmn && cmn_fuse_mode == 1) p.bulk_out = 0;   // mode 1: plain stores, the CTA re-reads its own rows after a block barrier
    if (fused_cmn && cmn_fuse_mode == 2 && !p.bulk_out) p.cmn_fused = 1;   // (bulk reductions need the bulk-store alignment)
    off += (size_t)p.smem_warp_stride * nw;
    if (off > 227 * 1024) return fail(MELSPEC_ERR_UNSUPPORTED, "hop_size too large for the shared-memory tile of this build");
    {   // MELSPEC_TILE_ORDER=0|1 overrides (A/B); default: interleaved warps for the mel-major layouts of plan 400
        static const int forced = [] { const char* e = std::getenv("MELSPEC_TILE_ORDER"); return e ? std::atoi(e) : -1; }();
        p.tile_order = forced >= 0 ? forced : (h->plan == 400 && layout == MELSPEC_LAYOUT_MEL_MAJOR ? 1 : 0);
        // measured (profiles/r2_ab_melmajor.txt): 80 rows per tile stay inside L2 without a barrier (and lose 4 - 19 % to one); 128 rows
        // do not: DRAM 1.54 GB -> 1.20 GB per launch and 0.61 -> 0.45 ms with a barrier every 8 passes
        static const int sync_forced = [] { const char* e = std::getenv("MELSPEC_MM_SYNC"); return e ? std::atoi(e) : -1; }();
        p.mm_sync = sync_forced >= 0 ? sync_forced : (c.n_mels > 80 ? 8 : 0);
    }
    const int64_t n_tiles = (n_wtiles + nw - 1) / nw;
    const int grid = fused_cmn ? h->num_sms : (int)std::min<int64_t>(n_tiles, h->num_sms);
    int32_t rc;
    const bool m3 = h->mpl <= 3;
    if (h->plan == 400) {
#define MS_DISPATCH(NW)                                                                                             \
    (h->kspec == 2 && hop160 ? launch_kernel(melspec400_kernel<NW, 3, true, 2>, p, grid, NW * 32, off, st)           \
     : h->kspec == 1 && hop160 ? launch_kernel(melspec400_kernel<NW, 3, true, 1>, p, grid, NW * 32, off, st)         \
     : m3 ? (hop160 ? launch_kernel(melspec400_kernel<NW, 3, true, 0>, p, grid, NW * 32, off, st)                    \
                    : launch_kernel(melspec400_kernel<NW, 3, false, 0>, p, grid, NW * 32, off, st))                   \
          : (hop160 ? launch_kernel(melspec400_kernel<NW, 4, true, 0>, p, grid, NW * 32, off, st)                    \
                    : launch_kernel(melspec400_kernel<NW, 4, false, 0>, p, grid, NW * 32, off, st)))
        // the launch shape of every large batch (frame-major, both TMA paths, no per-clip lengths) has its own instantiation with
        // those switches compiled in (KSPEC 3)
        const bool fshape = hop160 && nw == 12 && p.bulk_in && p.bulk_out && layout == MELSPEC_LAYOUT_FRAME_MAJOR && !d_lens && p.normalize;
        const bool fast = h->kspec == 1 && fshape;
        static const bool k4 = [] { const char* e = std::getenv("MELSPEC_KSPEC4"); return !(e && e[0] == '0'); }();   // (0: A/B against the table-driven loop)
        // ... and the mel-major layout of the 80-mel bank (interleave_frames / whisper.cpp images; MELSPEC_KSPEC5=0: A/B)
        static const bool k5 = [] { const char* e = std::getenv("MELSPEC_KSPEC5"); return !(e && e[0] == '0'); }();
        const bool fmm = k5 && h->kspec == 1 && c.n_mels == 80 && hop160 && nw == 12 && p.bulk_in && layout == MELSPEC_LAYOUT_MEL_MAJOR && p.mm_aligned8 &&
                         !d_lens && p.normalize && row_stride * 80 < ((int64_t)1 << 31);
        const bool fmm128 = k5 && h->kspec == 4 && c.n_mels == 128 && hop160 && nw == 12 && p.bulk_in && layout == MELSPEC_LAYOUT_MEL_MAJOR &&
                            p.mm_aligned8 && !d_lens && p.normalize && row_stride * 128 < ((int64_t)1 << 31);
        rc = fmm ? launch_kernel(melspec400_kernel<12, 3, true, 5>, p, grid, 12 * 32, off, st)
             : fmm128 ? launch_kernel(melspec400_kernel<12, 4, true, 6>, p, grid, 12 * 32, off, st)
             : h->kspec == 4 && fshape && k4 ? launch_kernel(melspec400_kernel<12, 4, true, 4>, p, grid, 12 * 32, off, st)
             : fast ? launch_kernel(melspec400_kernel<12, 3, true, 3>, p, grid, 12 * 32, off, st)
             : nw == 16 ? launch_kernel(melspec400_kernel<16, 3, true, 1>, p, grid, 16 * 32, off, st) : nw == 8 ? MS_DISPATCH(8) : MS_DISPATCH(12);
#undef MS_DISPATCH
    } else {
#define MS_DISPATCH(NW, MODE)                                                                                       \
    (m3 ? launch_kernel(melspec512_kernel<NW, 3, MODE>, p, grid, NW * 32, off, st)                                   \
        : launch_kernel(melspec512_kernel<NW, 4, MODE>, p, grid, NW * 32, off, st))
#define MS_DISPATCH_FAST(MODE)                                                                                      \
    (m3 ? launch_kernel(melspec512_kernel<12, 3, MODE, true>, p, grid, 12 * 32, off, st)                             \
        : launch_kernel(melspec512_kernel<12, 4, MODE, true>, p, grid, 12 * 32, off, st))
        // the launch shape of every large dense batch (aligned buffers: both TMA paths, no per-clip lengths) has its own
        // instantiation with those switches compiled in
        const bool fast = nw == 12 && p.bulk_in && !d_lens && (nemo || p.bulk_out || fused_cmn) && !(fused_cmn && p.cmn_fused == 1);
        // ... and the filterbanks of the frontends' defaults have their projection schedule compiled in (melspec::ksched512);
        // MELSPEC_KSCHED=0 falls back to the table-driven loop (A/B)
        static const bool ksched_on = [] { const char* e = std::getenv("MELSPEC_KSCHED"); return !(e && e[0] == '0'); }();
        const int ks = ksched_on ? h->ksched512 : 0;
        if (nemo) {
            // padding columns (pad_to) are zeros in the reference's feature matrix (src/mel.rs:336)
            if (row_stride > frames_per_clip && p.out_clip_stride == row_stride * c.n_mels) {
                const long long total = (long long)n_clips * c.n_mels * (row_stride - frames_per_clip);
                melspec_zero_cols_kernel<<<(int)std::min<long long>((total + 255) / 256, (long long)h->num_sms * 8), 256, 0, st>>>(
                    d_out, p.out_clip_stride, c.n_mels, row_stride, (int)frames_per_clip, (int)row_stride, n_clips);
                MS_CUDA(cudaGetLastError());
            }
            if (d_lens) rc = nw == 8 ? MS_DISPATCH(8, 3) : MS_DISPATCH(12, 3);   // ragged batch (per-clip lengths)
            else if (fast && ks == 3) rc = launch_kernel(melspec512_kernel<12, 3, 2, true, 3>, p, grid, 12 * 32, off, st);
            else if (fast && ks == 4) rc = launch_kernel(melspec512_kernel<12, 4, 2, true, 4>, p, grid, 12 * 32, off, st);
            else rc = fast ? MS_DISPATCH_FAST(2) : nw == 8 ? MS_DISPATCH(8, 2) : MS_DISPATCH(12, 2);
        } else if (kaldi) {
            if (fast && p.bulk_out && ks == 2) rc = launch_kernel(melspec512_kernel<12, 3, 1, true, 2>, p, grid, 12 * 32, off, st);
            else rc = fast && p.bulk_out ? MS_DISPATCH_FAST(1) : nw == 8 ? MS_DISPATCH(8, 1) : MS_DISPATCH(12, 1);
        } else {
            const bool f0 = fast && p.bulk_out && layout == MELSPEC_LAYOUT_FRAME_MAJOR;
            // Whisper-512 into the interleave_frames / whisper.cpp layout: the same compiled-in shape with the mel-major store
            const bool fmm = nw == 12 && p.bulk_in && !d_lens && ks == 1 && layout == MELSPEC_LAYOUT_MEL_MAJOR && p.normalize;
            if (fmm) rc = launch_kernel(melspec512_kernel<12, 3, 0, true, 1, true>, p, grid, 12 * 32, off, st);
            else if (f0 && ks == 1) rc = launch_kernel(melspec512_kernel<12, 3, 0, true, 1>, p, grid, 12 * 32, off, st);
            else if (f0 && ks == 4) rc = launch_kernel(melspec512_kernel<12, 4, 0, true, 4>, p, grid, 12 * 32, off, st);
            else rc = f0 ? MS_DISPATCH_FAST(0) : nw == 8 ? MS_DISPATCH(8, 0) : MS_DISPATCH(12, 0);
        }
#undef MS_DISPATCH
#undef MS_DISPATCH_FAST
    }
    if (rc != MELSPEC_OK) return rc;
    h->launches += 1;
    return launch_post_kernels(h, p, n_clips, d_lens, d_out, fused_cmn, st);
}

int32_t ensure_host_resources(melspec_handle* h, size_t pcm_bytes, size_t out_bytes, size_t i16_bytes = 0) {
    for (int i = 0; i < 3; ++i)
        if (!h->streams[i]) MS_CUDA(cudaStreamCreateWithFlags(&h->streams[i], cudaStreamNonBlocking));
    if (i16_bytes > h->slot_i16_cap) {
        for (int i = 0; i < 3; ++i) {
            if (h->d_slot_i16[i]) cudaFree(h->d_slot_i16[i]);
            h->d_slot_i16[i] = nullptr;
            MS_CUDA(cudaMalloc(&h->d_slot_i16[i], i16_bytes));
        }
        h->slot_i16_cap = i16_bytes;
    }
    if (pcm_bytes > h->slot_pcm_cap) {
        for (int i = 0; i < 3; ++i) {
            if (h->d_slot_pcm[i]) cudaFree(h->d_slot_pcm[i]);
            h->d_slot_pcm[i] = nullptr;
            MS_CUDA(cudaMalloc(&h->d_slot_pcm[i], pcm_bytes));
        }
        h->slot_pcm_cap = pcm_bytes;
    }
    if (out_bytes > h->slot_out_cap) {
        for (int i = 0; i < 3; ++i) {
            if (h->d_slot_out[i]) cudaFree(h->d_slot_out[i]);
            h->d_slot_out[i] = nullptr;
            MS_CUDA(cudaMalloc(&h->d_slot_out[i], out_bytes));
        }
        h->slot_out_cap = out_bytes;
    }
    return MELSPEC_OK;
}

}  // namespace

// ================================================================================================ C ABI
extern "C" {

int32_t melspec_abi_version(void) { return MELSPEC_B200_ABI_VERSION; }

const char* melspec_last_error(void) { return g_last_error.c_str(); }

int32_t melspec_default_config(int32_t frontend, melspec_config* cfg) {
    if (!cfg) return fail(MELSPEC_ERR_INVALID_ARG, "cfg is null");
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->frontend = frontend;
    cfg->hop_size = 160;
    cfg->n_mels = 80;
    cfg->sampling_rate = 16000.0;
    if (frontend == MELSPEC_FRONTEND_WHISPER) {
        cfg->fft_size = 400;
        cfg->frame_length = 400;
        return MELSPEC_OK;
    }
    if (frontend == MELSPEC_FRONTEND_KALDI) {   // FbankConfig::default(), src/fbank.rs:46-64
        cfg->frame_length = 400;
        cfg->fft_size = 512;
        cfg->apply_cmn = 1;
        cfg->use_log_fbank = 1;
        cfg->use_power = 1;
        cfg->preemphasis = 0.97;
        cfg->low_freq = 20.0;
        cfg->high_freq = 0.0;
        cfg->energy_floor = 0.0;
        return MELSPEC_OK;
    }
    if (frontend == MELSPEC_FRONTEND_NEMO) {    // BatchLogMelConfig::default(), src/mel.rs:189-208
        cfg->fft_size = 512;
        cfg->win_length = 400;
        cfg->frame_length = 400;
        cfg->center = 1;
        cfg->slaney_norm = 1;
        cfg->log_zero_guard = (double)1.1920928955078125e-07f;
        return MELSPEC_OK;
    }
    return fail(MELSPEC_ERR_INVALID_ARG, "unknown frontend");
}

int32_t melspec_build_filterbank(const melspec_config* cfg, double* out, int64_t capacity) {
    Resolved r;
    int32_t rc = resolve(cfg, r);
    if (rc) return rc;
    if (!out) return fail(MELSPEC_ERR_INVALID_ARG, "out is null");
    const int64_t need = (int64_t)r.n_mels * (r.fft / 2 + 1);
    if (capacity < need) return fail(MELSPEC_ERR_INVALID_ARG, "capacity too small for (n_mels, fft/2+1)");
    std::vector<double> w;
    build_filterbank(r, w);
    std::memcpy(out, w.data(), sizeof(double) * (size_t)need);
    return MELSPEC_OK;
}

int64_t melspec_num_frames_cfg(const melspec_config* cfg, int64_t n_samples) {
    Resolved r;
    if (resolve(cfg, r)) return -1;
    return frames_for(r, n_samples);
}

int32_t melspec_create(const melspec_config* cfg, int32_t device, melspec_handle** out) {
    if (!out) return fail(MELSPEC_ERR_INVALID_ARG, "out is null");
    *out = nullptr;
    Resolved r;
    int32_t rc = resolve(cfg, r);
    if (rc) return rc;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(MELSPEC_ERR_NO_DEVICE, "CUDA unavailable: no device (this library has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(MELSPEC_ERR_NO_DEVICE, "CUDA unavailable: device index out of range");
    cudaDeviceProp prop;
    MS_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(MELSPEC_ERR_NO_DEVICE, "CUDA unavailable: kernels are built for sm_100a (Blackwell) only");
    // Specialised kernels for the configurations the reference's tests / goldens / BASELINE.json name; the general plan
    // (melspec_generic.cuh) for every other size.  MELSPEC_FORCE_GENERIC=1 routes everything through the general plan (tests).
    static const bool force_generic = [] { const char* e = std::getenv("MELSPEC_FORCE_GENERIC"); return e && e[0] == '1'; }();
    int plan = 1;
    if (r.frontend == MELSPEC_FRONTEND_WHISPER && r.fft == 400 && r.hop <= 256) plan = 400;
    else if (r.frontend == MELSPEC_FRONTEND_WHISPER && r.fft == 512 && r.hop == 160) plan = 512;
    else if (r.frontend == MELSPEC_FRONTEND_KALDI && r.fft == 512 && r.frame_len == 400 && r.hop == 160 && r.use_power && r.use_log) plan = 512;
    else if (r.frontend == MELSPEC_FRONTEND_NEMO && r.fft == 512 && r.frame_len == 400 && r.hop == 160) plan = 512;
    if (force_generic) plan = 1;
    if (plan == 1) {
        // the CTA's twiddle table (8 N bytes) and one warp's two ping-pong buffers (16 N bytes, 8 N for even N) must fit
        const int nf = r.fft % 2 ? r.fft : r.fft / 2;
        if ((size_t)r.fft * 8 + (size_t)16 * melspec::generic_buf_elems(nf) > 220 * 1024)
            return fail(MELSPEC_ERR_UNSUPPORTED, "fft_size too large for the shared-memory FFT of this build (max 13652, odd sizes 9009)");
    }
    MS_CUDA(cudaSetDevice(device));
    melspec_handle* h = new (std::nothrow) melspec_handle();
    if (!h) return fail(MELSPEC_ERR_CUDA, "out of host memory");
    h->cfg = r;
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    h->plan = plan;
    build_filterbank(r, h->dense);
    if (h->plan != 1) {   // the specialised kernels never form bin 0: a bank with a non-zero DC column runs on the general plan
        const int nbins = r.fft / 2 + 1;
        for (int m = 0; m < r.n_mels; ++m)
            if (h->dense[(size_t)m * nbins] != 0.0) { h->plan = 1; break; }
    }
    rc = build_tables(h);
    if (rc) {
        melspec_destroy(h);
        return rc;
    }
    *out = h;
    return MELSPEC_OK;
}

void melspec_destroy(melspec_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    melspec_nccl_destroy(h);
    cudaFree(h->d_window);
    cudaFree(h->d_twiddle);
    cudaFree(h->d_rot10);
    cudaFree(h->d_proj);
    cudaFree(h->d_meta);
    cudaFree(h->d_gtw);
    cudaFree(h->d_gwin);
    cudaFree(h->d_gbands);
    cudaFree(h->d_gweights);
    cudaFree(h->d_gweights_t);
    cudaFree(h->d_gstarts_t);
    cudaFree(h->d_partials);
    for (int i = 0; i < 3; ++i) { cudaFree(h->d_slot_tga[i]); cudaFree(h->d_slot_part[i]); }
    cudaFree(h->d_fmt_img);
    cudaFree(h->d_fmt_tga);
    for (int i = 0; i < 3; ++i) {
        if (h->d_slot_pcm[i]) cudaFree(h->d_slot_pcm[i]);
        if (h->d_slot_out[i]) cudaFree(h->d_slot_out[i]);
        if (h->d_slot_i16[i]) cudaFree(h->d_slot_i16[i]);
        if (h->streams[i]) cudaStreamDestroy(h->streams[i]);
    }
    delete h;
}

int64_t melspec_num_frames(const melspec_handle* h, int64_t n_samples) { return h ? frames_for(h->cfg, n_samples) : -1; }
int64_t melspec_padded_frames(const melspec_handle* h, int64_t n_samples) { return h ? padded_frames_for(h->cfg, n_samples) : -1; }

int32_t melspec_max_frames_per_batch(const melspec_handle* h) {
    if (!h) return 0;
    // reference src/cuda.rs:150-155 with sizeof(cufftDoubleComplex) = 16, sizeof(f64) = 8
    const uint64_t per = 16ull * (uint64_t)h->cfg.fft + 8ull * (uint64_t)h->cfg.n_mels;
    const uint64_t v = std::max<uint64_t>((64ull * 1024 * 1024) / per, 1);
    return (int32_t)std::min<uint64_t>(v, 8192);
}

int32_t melspec_n_mels(const melspec_handle* h) { return h ? h->cfg.n_mels : 0; }
int32_t melspec_fft_size(const melspec_handle* h) { return h ? h->cfg.fft : 0; }
int32_t melspec_hop_size(const melspec_handle* h) { return h ? h->cfg.hop : 0; }

// ---- output formats: interleave_frames (src/mel.rs:480-544) and 8-bit TGA (src/quant.rs:38-165) -------------
int64_t melspec_interleaved_width(int64_t n_frames, int64_t min_width) {
    if (n_frames <= 0 || min_width < 0 || (min_width & 1)) return -1;   // asserts of src/mel.rs:487-488
    const int64_t even = n_frames + ((min_width > 0 && (n_frames & 1)) ? 1 : 0);   // src/mel.rs:497-500
    return std::max(even, min_width);                                               // src/mel.rs:506-516
}

int32_t melspec_compute_interleaved_device(melspec_handle* h, const float* d_pcm, int64_t n_clips, int64_t clip_stride,
                                           int64_t n_samples, int64_t min_width, float* d_out, int64_t out_clip_stride, void* stream) {
    if (!h) return fail(MELSPEC_ERR_INVALID_ARG, "handle is null");
    if (h->cfg.frontend != MELSPEC_FRONTEND_WHISPER)
        return fail(MELSPEC_ERR_UNSUPPORTED, "interleave_frames is defined on Whisper mel frames (src/mel.rs:480-544)");
    if (n_clips < 0 || n_samples < 0 || clip_stride < 0 || out_clip_stride < 0) return fail(MELSPEC_ERR_INVALID_ARG, "negative size");
    if (min_width < 0 || (min_width & 1)) return fail(MELSPEC_ERR_INVALID_ARG, "min_width must be even");           // src/mel.rs:488
    const int64_t F = frames_for(h->cfg, n_samples);
    if (F == 0) return fail(MELSPEC_ERR_INVALID_ARG, "frames is empty");                                            // src/mel.rs:487
    if (n_clips == 0) return MELSPEC_OK;
    if (!d_pcm || !d_out) return fail(MELSPEC_ERR_INVALID_ARG, "null device pointer");
    if (n_clips > 1 && clip_stride < n_samples) return fail(MELSPEC_ERR_INVALID_ARG, "clip_stride < n_samples");
    const int64_t W = melspec_interleaved_width(F, min_width);
    const int64_t ocs = out_clip_stride ? out_clip_stride : W * h->cfg.n_mels;
    if (ocs < W * h->cfg.n_mels) return fail(MELSPEC_ERR_INVALID_ARG, "out_clip_stride < n_mels * width");
    MS_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (W > F) {   // the zero frame / zero padding block of src/mel.rs:497-516
        const long long total = (long long)n_clips * h->cfg.n_mels * (W - F);
        const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)h->num_sms * 8);
        melspec::melspec_zero_cols_kernel<<<blocks, 256, 0, st>>>(d_out, ocs, h->cfg.n_mels, W, (int)F, (int)W, n_clips);
        MS_CUDA(cudaGetLastError());
        h->launches += 1;
    }
    return launch_device(h, d_pcm, n_clips, clip_stride, n_samples, F, nullptr, d_out, ocs, MELSPEC_LAYOUT_MEL_MAJOR, st, W);
}

int64_t melspec_tga_size(int32_t n_mels, int64_t width) {
    if (n_mels <= 0 || width <= 0 || n_mels > 65535 || width > 65535) return -1;   // u16 header fields (src/quant.rs:44-57)
    return (int64_t)melspec::kTgaHeader + (int64_t)n_mels * width;
}

namespace {
int quantize_partials_per_image(int64_t n) { return (int)std::min<int64_t>(256, (n + 4095) / 4096); }
// min/max partials + quantise, with the partials workspace given by the caller (one per stream that may be in flight)
int32_t quantize_tga_launch(melspec_handle* h, const float* d_img, int64_t n_imgs, int64_t img_stride, int32_t n_mels, int64_t width,
                            uint8_t* d_tga, int64_t tga_stride, float2* d_partials, cudaStream_t st) {
    const int64_t n = (int64_t)n_mels * width;
    const int nblk = quantize_partials_per_image(n);
    melspec::melspec_minmax_kernel<<<dim3(nblk, (unsigned)n_imgs), 256, 0, st>>>(d_img, img_stride, n, d_partials);
    MS_CUDA(cudaGetLastError());
    const int nblk2 = (int)std::min<int64_t>(1024, (n / 4 + 1023) / 1024 + 1);
    melspec::melspec_quantize_kernel<<<dim3(nblk2, (unsigned)n_imgs), 256, 0, st>>>(d_img, img_stride, n, d_partials, nblk, d_tga, tga_stride,
                                                                                  n_mels, (int)width);
    MS_CUDA(cudaGetLastError());
    h->launches += 2;
    return MELSPEC_OK;
}
}  // namespace

int32_t melspec_quantize_tga_device(melspec_handle* h, const float* d_img, int64_t n_imgs, int64_t img_stride, int32_t n_mels,
                                    int64_t width, uint8_t* d_tga, int64_t tga_stride, void* stream) {
    if (!h) return fail(MELSPEC_ERR_INVALID_ARG, "handle is null");
    const int64_t sz = melspec_tga_size(n_mels, width);
    if (sz < 0) return fail(MELSPEC_ERR_INVALID_ARG, "width greater than TARGA max (or empty image)");   // src/quant.rs:18-21
    if (n_imgs < 0 || img_stride < 0 || tga_stride < 0) return fail(MELSPEC_ERR_INVALID_ARG, "negative size");
    if (n_imgs == 0) return MELSPEC_OK;
    if (!d_img || !d_tga) return fail(MELSPEC_ERR_INVALID_ARG, "null device pointer");
    const int64_t n = (int64_t)n_mels * width;
    if (!img_stride) img_stride = n;
    if (!tga_stride) tga_stride = sz;
    if (n_imgs > 1 && (img_stride < n || tga_stride < sz)) return fail(MELSPEC_ERR_INVALID_ARG, "stride smaller than the image");
    if (n_imgs > 65535) return fail(MELSPEC_ERR_INVALID_ARG, "too many images for one call");
    MS_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t need = sizeof(float2) * (size_t)quantize_partials_per_image(n) * (size_t)n_imgs;
    if (need > h->partials_cap) {
        if (h->d_partials) { MS_CUDA(cudaStreamSynchronize(st)); cudaFree(h->d_partials); h->d_partials = nullptr; }
        MS_CUDA(cudaMalloc(&h->d_partials, need));
        h->partials_cap = need;
    }
    return quantize_tga_launch(h, d_img, n_imgs, img_stride, n_mels, width, d_tga, tga_stride, h->d_partials, st);
}

int32_t melspec_dequantize_tga_device(melspec_handle* h, const uint8_t* d_tga, int64_t n_imgs, int64_t tga_stride, int32_t n_mels,
                                      int64_t width, float* d_img, int64_t img_stride, void* stream) {
    if (!h) return fail(MELSPEC_ERR_INVALID_ARG, "handle is null");
    const int64_t sz = melspec_tga_size(n_mels, width);
    if (sz < 0) return fail(MELSPEC_ERR_INVALID_ARG, "width greater than TARGA max (or empty image)");
    if (n_imgs < 0 || img_stride < 0 || tga_stride < 0) return fail(MELSPEC_ERR_INVALID_ARG, "negative size");
    if (n_imgs == 0) return MELSPEC_OK;
    if (!d_img || !d_tga) return fail(MELSPEC_ERR_INVALID_ARG, "null device pointer");
    const int64_t n = (int64_t)n_mels * width;
    if (!img_stride) img_stride = n;
    if (!tga_stride) tga_stride = sz;
    if (n_imgs > 65535) return fail(MELSPEC_ERR_INVALID_ARG, "too many images for one call");
    MS_CUDA(cudaSetDevice(h->device));
    const int nblk = (int)std::min<int64_t>(1024, (n + 8191) / 8192);   // 256 threads x 8 quads of 4 pixels
    melspec::melspec_dequantize_kernel<<<dim3(nblk, (unsigned)n_imgs), 256, 0, (cudaStream_t)stream>>>(d_tga, tga_stride, n, d_img, img_stride);
    MS_CUDA(cudaGetLastError());
    h->launches += 1;
    return MELSPEC_OK;
}

namespace {
int32_t ensure_fmt(melspec_handle* h, size_t img_bytes, size_t tga_bytes) {
    if (img_bytes > h->fmt_img_cap) {
        if (h->d_fmt_img) cudaFree(h->d_fmt_img);
        h->d_fmt_img = nullptr; h->fmt_img_cap = 0;
        MS_CUDA(cudaMalloc(&h->d_fmt_img, img_bytes));
        h->fmt_img_cap = img_bytes;
    }
    if (tga_bytes > h->fmt_tga_cap) {
        if (h->d_fmt_tga) cudaFree(h->d_fmt_tga);
        h->d_fmt_tga = nullptr; h->fmt_tga_cap = 0;
        MS_CUDA(cudaMalloc(&h->d_fmt_tga, tga_bytes));
        h->fmt_tga_cap = tga_bytes;
    }
    return MELSPEC_OK;
}
}  // namespace

// Host-buffer conveniences (blocking).  h_img is row-major (n_mels, width) as produced by interleave_frames(.., false, w).
int32_t melspec_quantize_tga_host(melspec_handle* h, const float* h_img, int32_t n_mels, int64_t width, uint8_t* h_tga) {
    if (!h) return fail(MELSPEC_ERR_INVALID_ARG, "handle is null");
    const int64_t sz = melspec_tga_size(n_mels, width);
    if (sz < 0) return fail(MELSPEC_ERR_INVALID_ARG, "width greater than TARGA max (or empty image)");
    if (!h_img || !h_tga) return fail(MELSPEC_ERR_INVALID_ARG, "null host pointer");
    MS_CUDA(cudaSetDevice(h->device));
    const size_t ib = (size_t)n_mels * (size_t)width * 4;
    int32_t rc = ensure_fmt(h, ib, (size_t)sz + 8);
    if (rc) return rc;
    MS_CUDA(cudaMemcpy(h->d_fmt_img, h_img, ib, cudaMemcpyHostToDevice));
    rc = melspec_quantize_tga_device(h, h->d_fmt_img, 1, 0, n_mels, width, h->d_fmt_tga, 0, nullptr);
    if (rc) return rc;
    MS_CUDA(cudaMemcpy(h_tga, h->d_fmt_tga, (size_t)sz, cudaMemcpyDeviceToHost));
    return MELSPEC_OK;
}

// parse_tga_8bit (src/quant.rs:66-88): width and height come from the caller (the reference ignores the header's, too).
int32_t melspec_dequantize_tga_host(melspec_handle* h, const uint8_t* h_tga, int64_t tga_bytes, float* h_img, int64_t capacity) {
    if (!h) return fail(MELSPEC_ERR_INVALID_ARG, "handle is null");
    if (!h_tga || !h_img) return fail(MELSPEC_ERR_INVALID_ARG, "null host pointer");
    if (tga_bytes < melspec::kTgaHeader) return fail(MELSPEC_ERR_INVALID_ARG, "failed to fill whole buffer");   // read_exact error of the reference
    const int64_t n = tga_bytes - melspec::kTgaHeader;
    if (capacity < n) return fail(MELSPEC_ERR_INVALID_ARG, "capacity too small");
    if (n == 0) return MELSPEC_OK;
    MS_CUDA(cudaSetDevice(h->device));
    int32_t rc = ensure_fmt(h, (size_t)n * 4, (size_t)tga_bytes + 8);
    if (rc) return rc;
    MS_CUDA(cudaMemcpy(h->d_fmt_tga, h_tga, (size_t)tga_bytes, cudaMemcpyHostToDevice));
    const int nblk = (int)std::min<int64_t>(1024, (n + 8191) / 8192);   // 256 threads x 8 quads of 4 pixels
    melspec::melspec_dequantize_kernel<<<dim3(nblk, 1), 256>>>(h->d_fmt_tga, tga_bytes, n, h->d_fmt_img, n);
    MS_CUDA(cudaGetLastError());
    h->launches += 1;
    MS_CUDA(cudaMemcpy(h_img, h->d_fmt_img, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return MELSPEC_OK;
}

// PCM -> mel -> interleave -> 8-bit TGA in one call, everything between the two copies on the device
// (the reference's examples/mel_tga pipeline: src/stft.rs + src/mel.rs:480-544 + src/quant.rs:38-64).
int32_t melspec_mel_tga_host(melspec_handle* h, const float* h_pcm, int64_t n_samples, int64_t min_width, uint8_t* h_tga,
                             int64_t capacity, int64_t* width_out, float* h_img_opt) {
    if (!h) return fail(MELSPEC_ERR_INVALID_ARG, "handle is null");
    if (h->cfg.frontend != MELSPEC_FRONTEND_WHISPER) return fail(MELSPEC_ERR_UNSUPPORTED, "Whisper frontend only");
    if (n_samples < 0) return fail(MELSPEC_ERR_INVALID_ARG, "negative size");
    if (min_width < 0 || (min_width & 1)) return fail(MELSPEC_ERR_INVALID_ARG, "min_width must be even");
    const int64_t F = frames_for(h->cfg, n_samples);
    if (F == 0) return fail(MELSPEC_ERR_INVALID_ARG, "frames is empty");
    const int64_t W = melspec_interleaved_width(F, min_width);
    if (width_out) *width_out = W;
    const int64_t sz = melspec_tga_size(h->cfg.n_mels, W);
    if (sz < 0) return fail(MELSPEC_ERR_INVALID_ARG, "width greater than TARGA max, use chunks (src/quant.rs:17-21)");
    if (!h_pcm || !h_tga) return fail(MELSPEC_ERR_INVALID_ARG, "null host pointer");
    if (capacity < sz) return fail(MELSPEC_ERR_INVALID_ARG, "capacity too small");
    MS_CUDA(cudaSetDevice(h->device));
    const int64_t ns4 = (n_samples + 3) / 4 * 4;
    int32_t rc = ensure_host_resources(h, (size_t)ns4 * 4, 16);
    if (rc) return rc;
    rc = ensure_fmt(h, (size_t)h->cfg.n_mels * (size_t)W * 4, (size_t)sz + 8);
    if (rc) return rc;
    cudaStream_t st = h->streams[0];
    MS_CUDA(cudaMemcpyAsync(h->d_slot_pcm[0], h_pcm, (size_t)n_samples * 4, cudaMemcpyHostToDevice, st));
    rc = melspec_compute_interleaved_device(h, h->d_slot_pcm[0], 1, ns4, n_samples, min_width, h->d_fmt_img, 0, st);
    if (rc) return rc;
    rc = melspec_quantize_tga_device(h, h->d_fmt_img, 1, 0, h->cfg.n_mels, W, h->d_fmt_tga, 0, st);
    if (rc) return rc;
    MS_CUDA(cudaMemcpyAsync(h_tga, h->d_fmt_tga, (size_t)sz, cudaMemcpyDeviceToHost, st));
    if (h_img_opt) MS_CUDA(cudaMemcpyAsync(h_img_opt, h->d_fmt_img, (size_t)h->cfg.n_mels * (size_t)W * 4, cudaMemcpyDeviceToHost, st));
    MS_CUDA(cudaStreamSynchronize(st));
    return MELSPEC_OK;
}


// ---- VAD over the mel image (reference src/vad.rs:251-486, 163-207) -----------------------------------------
int32_t melspec_vad_default_settings(melspec_vad_settings* s) {
    if (!s) return fail(MELSPEC_ERR_INVALID_ARG, "settings is null");
    s->min_energy = 0.98; s->min_y = 11; s->min_x = 5; s->min_mel = 2;   // src/vad.rs:13-22
    return MELSPEC_OK;
}

int32_t melspec_vad_boundaries_device(melspec_handle* h, const float* d_img, int64_t n_imgs, int64_t img_stride, int32_t n_mels,
                                      int64_t width, const melspec_vad_settings* vs, uint8_t* d_raw, uint8_t* d_smoothed,
                                      int64_t mask_stride, void* stream) {
    if (!h) return fail(MELSPEC_ERR_INVALID_ARG, "handle is null");
    if (!vs) return fail(MELSPEC_ERR_INVALID_ARG, "settings is null");
    if (n_imgs < 0 || img_stride < 0 || mask_stride < 0 || n_mels < 0 || width < 0) return fail(MELSPEC_ERR_INVALID_ARG, "negative size");
    if (vs->min_y < 0 || vs->min_x < 0 || vs->min_mel < 0) return fail(MELSPEC_ERR_INVALID_ARG, "negative setting");
    if (width > 0x7fffffff || (int64_t)n_mels * width > 0x7fffffffll * 4) return fail(MELSPEC_ERR_INVALID_ARG, "image too large");
    if (n_imgs == 0 || n_mels < 3 || width < 3) return MELSPEC_OK;        // empty EdgeInfo, src/vad.rs:264-266
    if (!d_img || !d_smoothed) return fail(MELSPEC_ERR_INVALID_ARG, "null device pointer");
    if (n_imgs > 65535) return fail(MELSPEC_ERR_INVALID_ARG, "too many images for one call");
    if (!img_stride) img_stride = (int64_t)n_mels * width;
    if (!mask_stride) mask_stride = width - 2;
    MS_CUDA(cudaSetDevice(h->device));
    const int nb = (int)((width - 2 + melspec::kVadTile - 1) / melspec::kVadTile);
    melspec::melspec_vad_kernel<<<dim3(nb, (unsigned)n_imgs), 256, 0, (cudaStream_t)stream>>>(
        d_img, img_stride, n_mels, (int)width, vs->min_energy * vs->min_energy, vs->min_y, vs->min_mel, d_raw, d_smoothed, mask_stride);
    MS_CUDA(cudaGetLastError());
    h->launches += 1;
    return MELSPEC_OK;
}

int32_t melspec_vad_activity_device(melspec_handle* h, const uint8_t* d_raw, int64_t n_imgs, int64_t mask_stride, int32_t n_mels,
                                    int64_t width, const melspec_vad_settings* vs, int32_t* d_activity, int64_t activity_stride,
                                    void* stream) {
    if (!h) return fail(MELSPEC_ERR_INVALID_ARG, "handle is null");
    if (!vs) return fail(MELSPEC_ERR_INVALID_ARG, "settings is null");
    if (n_imgs < 0 || mask_stride < 0 || activity_stride < 0 || width < 0 || width > 0x7fffffff) return fail(MELSPEC_ERR_INVALID_ARG, "bad size");
    if (vs->min_x < 0) return fail(MELSPEC_ERR_INVALID_ARG, "negative setting");
    if (n_imgs == 0 || width == 0) return MELSPEC_OK;
    if (!d_activity || (!d_raw && n_mels >= 3 && width >= 3)) return fail(MELSPEC_ERR_INVALID_ARG, "null device pointer");
    if (n_imgs > 65535) return fail(MELSPEC_ERR_INVALID_ARG, "too many images for one call");
    if (!mask_stride) mask_stride = width - 2;
    if (!activity_stride) activity_stride = 3 * width;
    MS_CUDA(cudaSetDevice(h->device));
    melspec::melspec_vad_activity_kernel<<<dim3((unsigned)((width + 255) / 256), (unsigned)n_imgs), 256, 0, (cudaStream_t)stream>>>(
        d_raw, mask_stride, n_mels, (int)width, vs->min_x, d_activity, activity_stride);
    MS_CUDA(cudaGetLastError());
    h->launches += 1;
    return MELSPEC_OK;
}

// Host-buffer convenience (blocking): one row-major (n_mels, width) f32 image -> smoothed mask (width-2 bytes, 1 = the column
// intersects an edge) and, optionally, the per-frame activity triples (3*width int32).
int32_t melspec_vad_host(melspec_handle* h, const float* h_img, int32_t n_mels, int64_t width, const melspec_vad_settings* vs,
                         uint8_t* h_smoothed, int32_t* h_activity_opt) {
    if (!h) return fail(MELSPEC_ERR_INVALID_ARG, "handle is null");
    if (!vs) return fail(MELSPEC_ERR_INVALID_ARG, "settings is null");
    if (n_mels < 0 || width < 0) return fail(MELSPEC_ERR_INVALID_ARG, "negative size");
    if (width == 0 || n_mels == 0) return MELSPEC_OK;
    if (!h_img) return fail(MELSPEC_ERR_INVALID_ARG, "null host pointer");
    MS_CUDA(cudaSetDevice(h->device));
    const size_t ib = (size_t)n_mels * (size_t)width * 4;
    const size_t nmask = width >= 3 ? (size_t)(width - 2) : 0;
    const size_t act_off = (2 * nmask + 15) / 16 * 16;
    int32_t rc = ensure_fmt(h, ib, act_off + 12 * (size_t)width + 16);
    if (rc) return rc;
    MS_CUDA(cudaMemcpy(h->d_fmt_img, h_img, ib, cudaMemcpyHostToDevice));
    uint8_t* d_raw = h->d_fmt_tga;
    uint8_t* d_sm = h->d_fmt_tga + nmask;
    int32_t* d_act = reinterpret_cast<int32_t*>(h->d_fmt_tga + act_off);
    if (nmask) MS_CUDA(cudaMemset(d_raw, 0, 2 * nmask));
    rc = melspec_vad_boundaries_device(h, h->d_fmt_img, 1, 0, n_mels, width, vs, d_raw, d_sm, 0, nullptr);
    if (rc) return rc;
    if (h_smoothed && nmask && n_mels >= 3) MS_CUDA(cudaMemcpy(h_smoothed, d_sm, nmask, cudaMemcpyDeviceToHost));
    if (h_activity_opt) {
        rc = melspec_vad_activity_device(h, d_raw, 1, 0, n_mels, width, vs, d_act, 0, nullptr);
        if (rc) return rc;
        MS_CUDA(cudaMemcpy(h_activity_opt, d_act, 12 * (size_t)width, cudaMemcpyDeviceToHost));
    }
    return MELSPEC_OK;
}

// ---- optional NCCL gather of the output shards: libnccl.so.2 resolved at run time -------------------------
namespace {
struct NcclId128 { char b[128]; };   // ncclUniqueId (passed by value to ncclCommInitRank)
struct NcclApi {
    int (*get_unique_id)(void*) = nullptr;
    int (*comm_init_rank)(void**, int, NcclId128, int) = nullptr;
    int (*all_gather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*comm_destroy)(void*) = nullptr;
    const char* (*error_string)(int) = nullptr;
    bool ok = false;
};
NcclApi& nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy the process already has (PyTorch's), if any
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return a;
        a.get_unique_id = reinterpret_cast<decltype(a.get_unique_id)>(dlsym(lib, "ncclGetUniqueId"));
        a.comm_init_rank = reinterpret_cast<decltype(a.comm_init_rank)>(dlsym(lib, "ncclCommInitRank"));
        a.all_gather = reinterpret_cast<decltype(a.all_gather)>(dlsym(lib, "ncclAllGather"));
        a.comm_destroy = reinterpret_cast<decltype(a.comm_destroy)>(dlsym(lib, "ncclCommDestroy"));
        a.error_string = reinterpret_cast<decltype(a.error_string)>(dlsym(lib, "ncclGetErrorString"));
        a.ok = a.get_unique_id && a.comm_init_rank && a.all_gather && a.comm_destroy && a.error_string;
        return a;
    }();
    return api;
}
int32_t fail_nccl(int rc, const char* what) {
    return fail(MELSPEC_ERR_CUDA, std::string(what) + ": " + nccl_api().error_string(rc));
}
}  // namespace

int32_t melspec_nccl_unique_id(uint8_t* id128) {
    if (!id128) return fail(MELSPEC_ERR_INVALID_ARG, "id is null");
    if (!nccl_api().ok) return fail(MELSPEC_ERR_UNSUPPORTED, "libnccl.so.2 not found (the gather is optional; the hot path has no collective)");
    const int rc = nccl_api().get_unique_id(id128);
    return rc ? fail_nccl(rc, "ncclGetUniqueId") : MELSPEC_OK;
}

int32_t melspec_nccl_init(melspec_handle* h, const uint8_t* id128, int32_t rank, int32_t world_size) {
    if (!h || !id128) return fail(MELSPEC_ERR_INVALID_ARG, "null argument");
    if (world_size < 1 || rank < 0 || rank >= world_size) return fail(MELSPEC_ERR_INVALID_ARG, "rank outside [0, world_size)");
    if (!nccl_api().ok) return fail(MELSPEC_ERR_UNSUPPORTED, "libnccl.so.2 not found (the gather is optional; the hot path has no collective)");
    if (h->nccl_comm) return fail(MELSPEC_ERR_INVALID_ARG, "communicator already initialised for this handle");
    MS_CUDA(cudaSetDevice(h->device));
    NcclId128 id;
    std::memcpy(id.b, id128, 128);
    const int rc = nccl_api().comm_init_rank(&h->nccl_comm, world_size, id, rank);
    if (rc) { h->nccl_comm = nullptr; return fail_nccl(rc, "ncclCommInitRank"); }
    h->nccl_world = world_size;
    return MELSPEC_OK;
}

int32_t melspec_gather_nccl(melspec_handle* h, const float* d_shard, int64_t count, float* d_full, void* stream) {
    if (!h) return fail(MELSPEC_ERR_INVALID_ARG, "handle is null");
    if (!h->nccl_comm) return fail(MELSPEC_ERR_INVALID_ARG, "melspec_nccl_init has not been called on this handle");
    if (count < 0) return fail(MELSPEC_ERR_INVALID_ARG, "negative count");
    if (count == 0) return MELSPEC_OK;
    if (!d_shard || !d_full) return fail(MELSPEC_ERR_INVALID_ARG, "null device pointer");
    MS_CUDA(cudaSetDevice(h->device));
    const int rc = nccl_api().all_gather(d_shard, d_full, (size_t)count, /* ncclFloat32 */ 7, h->nccl_comm, (cudaStream_t)stream);
    return rc ? fail_nccl(rc, "ncclAllGather") : MELSPEC_OK;
}

int32_t melspec_nccl_destroy(melspec_handle* h) {
    if (!h) return fail(MELSPEC_ERR_INVALID_ARG, "handle is null");
    if (h->nccl_comm) {
        cudaSetDevice(h->device);
        nccl_api().comm_destroy(h->nccl_comm);
        h->nccl_comm = nullptr;
        h->nccl_world = 0;
    }
    return MELSPEC_OK;
}

int64_t melspec_launch_count(const melspec_handle* h) { return h ? h->launches : 0; }

int32_t melspec_filterbank(const melspec_handle* h, double* out, int64_t capacity) {
    if (!h || !out) return fail(MELSPEC_ERR_INVALID_ARG, "null argument");
    if (capacity < (int64_t)h->dense.size()) return fail(MELSPEC_ERR_INVALID_ARG, "capacity too small");
    std::memcpy(out, h->dense.data(), sizeof(double) * h->dense.size());
    return MELSPEC_OK;
}

int32_t melspec_compute_device(melspec_handle* h, const float* d_pcm, int64_t n_clips, int64_t clip_stride,
                               int64_t n_samples, const int32_t* d_lens, float* d_out, int64_t out_clip_stride,
                               int32_t layout, void* stream) {
    if (!h) return fail(MELSPEC_ERR_INVALID_ARG, "handle is null");
    if (n_clips < 0 || n_samples < 0 || clip_stride < 0 || out_clip_stride < 0)
        return fail(MELSPEC_ERR_INVALID_ARG, "negative size");
    if (layout != MELSPEC_LAYOUT_FRAME_MAJOR && layout != MELSPEC_LAYOUT_MEL_MAJOR)
        return fail(MELSPEC_ERR_INVALID_ARG, "unknown layout");
    const int64_t F = frames_for(h->cfg, n_samples);
    if (n_clips == 0 || F == 0) return MELSPEC_OK;   // empty input => no frames, success (src/cuda.rs:91-93)
    if (!d_pcm || !d_out) return fail(MELSPEC_ERR_INVALID_ARG, "null device pointer");
    if (n_clips > 1 && clip_stride < n_samples) return fail(MELSPEC_ERR_INVALID_ARG, "clip_stride < n_samples");
    MS_CUDA(cudaSetDevice(h->device));
    return launch_device(h, d_pcm, n_clips, clip_stride, n_samples, F, d_lens, d_out, out_clip_stride, layout, (cudaStream_t)stream);
}

namespace {
// int16 rows on the device -> f32 rows (x / 32768), asynchronous on `st`
int32_t launch_convert_i16(melspec_handle* h, const int16_t* d_in, int64_t n_rows, int64_t in_stride, int64_t n, float* d_out,
                           int64_t out_stride, cudaStream_t st) {
    if (n_rows == 0 || n == 0) return MELSPEC_OK;
    const int vec = ((uintptr_t)d_in % 16 == 0) && ((uintptr_t)d_out % 16 == 0) && (in_stride % 8 == 0) && (out_stride % 4 == 0);
    int64_t done = 0;
    while (done < n_rows) {   // gridDim.y <= 65535
        const int64_t rows = std::min<int64_t>(n_rows - done, 65535);
        const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n / 8 + 255) / 256, std::max<int64_t>(1, 8 * h->num_sms / rows)));
        melspec::melspec_i16_to_f32_kernel<<<dim3(blocks, (unsigned)rows), 256, 0, st>>>(d_in + done * in_stride, in_stride, (int)n,
                                                                                   d_out + done * out_stride, out_stride, vec);
        MS_CUDA(cudaGetLastError());
        h->launches += 1;
        done += rows;
    }
    return MELSPEC_OK;
}

// H2D of `nc` rows of `ns` samples (host row stride `clip_stride`, first sample `h_off` of each row) into staging slot `slot`
// (device row stride `dstride`): f32 directly, int16 through the int16 slot and the conversion kernel.
int32_t stage_rows(melspec_handle* h, int slot, const void* h_pcm, bool i16, int64_t row0, int64_t nc, int64_t clip_stride,
                   int64_t h_off, int64_t ns, int64_t dstride, cudaStream_t st) {
    const size_t es = i16 ? 2 : 4;
    const char* src = static_cast<const char*>(h_pcm) + (size_t)(row0 * clip_stride + h_off) * es;
    void* dst = i16 ? static_cast<void*>(h->d_slot_i16[slot]) : static_cast<void*>(h->d_slot_pcm[slot]);
    if (clip_stride == dstride || nc == 1) {
        MS_CUDA(cudaMemcpyAsync(dst, src, (size_t)((nc - 1) * dstride + ns) * es, cudaMemcpyHostToDevice, st));
    } else {
        MS_CUDA(cudaMemcpy2DAsync(dst, (size_t)dstride * es, src, (size_t)clip_stride * es, (size_t)ns * es, (size_t)nc,
                                  cudaMemcpyHostToDevice, st));
    }
    if (i16) return launch_convert_i16(h, h->d_slot_i16[slot], nc, dstride, ns, h->d_slot_pcm[slot], dstride, st);
    return MELSPEC_OK;
}

int32_t compute_host_impl(melspec_handle* h, const void* h_pcm, bool i16, int64_t n_clips, int64_t clip_stride, int64_t n_samples,
                          float* h_out, int32_t layout, int64_t* frames_out) {
    if (!h) return fail(MELSPEC_ERR_INVALID_ARG, "handle is null");
    if (n_clips < 0 || n_samples < 0 || clip_stride < 0) return fail(MELSPEC_ERR_INVALID_ARG, "negative size");
    if (layout != MELSPEC_LAYOUT_FRAME_MAJOR && layout != MELSPEC_LAYOUT_MEL_MAJOR)
        return fail(MELSPEC_ERR_INVALID_ARG, "unknown layout");
    const int64_t F = frames_for(h->cfg, n_samples);
    if (frames_out) *frames_out = F;
    if (n_clips == 0 || F == 0) return MELSPEC_OK;
    if (!h_pcm || !h_out) return fail(MELSPEC_ERR_INVALID_ARG, "null host pointer");
    if (n_clips > 1 && clip_stride < n_samples) return fail(MELSPEC_ERR_INVALID_ARG, "clip_stride < n_samples");
    MS_CUDA(cudaSetDevice(h->device));
    // Clips are cut into chunks that rotate over 3 (stream, device slot) pairs, so the H2D copy of chunk i+1, the
    // kernel of chunk i and the D2H copy of chunk i-1 overlap when the host buffers are pinned.
    const int64_t ns4 = (n_samples + 7) / 8 * 8;   // device rows are padded to 32 bytes so the TMA path (and the int16 vector loads) apply
    const int64_t clip_out = padded_frames_for(h->cfg, n_samples) * h->cfg.n_mels;
    static const int64_t chunk_mb = [] {            // MELSPEC_HOST_CHUNK_MB: PCM bytes per pipeline chunk (tuning knob)
        const char* e = std::getenv("MELSPEC_HOST_CHUNK_MB");
        const int64_t v = e ? std::atoll(e) : 0;
        return v > 0 ? v : (int64_t)32;
    }();
    const int64_t target = chunk_mb << 20;
    // One long clip (the reference's own call shape, `compute_mel_spectrogram(&[f32])`, on minutes to hours of audio): cut it
    // along time into pieces of whole warp tiles, so that the H2D copy of piece i+1, the kernel of piece i and the D2H copy
    // of piece i-1 overlap like the chunks of a batch do.  Whisper frames depend on their own samples only, so a piece is
    // just a shorter clip starting at frame f0 (its samples [f0 hop, (f1-1) hop + N) are re-read with their 240-sample
    // halo); pieces are multiples of 12 frames, which keeps every frame in the same slot of its transform as in the
    // unsplit launch: the result is bit-identical.  Kaldi (CMN, look-back) and NeMo (centre padding) are not cut.
    const int64_t piece_bytes = std::min<int64_t>(target, 8 << 20);   // 8 MB pieces: short ramp-up and tail, still DMA-efficient
    if (h->cfg.frontend == MELSPEC_FRONTEND_WHISPER && n_clips == 1 && n_samples * 4 >= 2 * piece_bytes) {
        const Resolved& c = h->cfg;
        int64_t pf = std::max<int64_t>(12, piece_bytes / 4 / c.hop / 12 * 12);   // frames per piece
        const int64_t psamples = ((pf - 1) * c.hop + c.fft + 7) / 8 * 8;
        int32_t rc = ensure_host_resources(h, (size_t)psamples * 4, (size_t)pf * c.n_mels * 4, i16 ? (size_t)psamples * 2 : 0);
        if (rc) return rc;
        int slot = 0;
        for (int64_t f0 = 0; f0 < F; f0 += pf, slot = (slot + 1) % 3) {
            const int64_t nf = std::min(pf, F - f0);
            const int64_t s0 = f0 * c.hop, ns = (nf - 1) * c.hop + c.fft;
            cudaStream_t st = h->streams[slot];
            rc = stage_rows(h, slot, h_pcm, i16, 0, 1, psamples, s0, ns, psamples, st);
            if (rc) return rc;
            rc = launch_device(h, h->d_slot_pcm[slot], 1, psamples, ns, nf, nullptr, h->d_slot_out[slot], 0, layout, st);
            if (rc) return rc;
            if (layout == MELSPEC_LAYOUT_FRAME_MAJOR)
                MS_CUDA(cudaMemcpyAsync(h_out + f0 * c.n_mels, h->d_slot_out[slot], (size_t)nf * c.n_mels * 4, cudaMemcpyDeviceToHost, st));
            else   // mel-major: the piece's [n_mels][nf] block goes into columns [f0, f0 + nf) of the [n_mels][F] image
                MS_CUDA(cudaMemcpy2DAsync(h_out + f0, (size_t)F * 4, h->d_slot_out[slot], (size_t)nf * 4, (size_t)nf * 4,
                                          (size_t)c.n_mels, cudaMemcpyDeviceToHost, st));
        }
        for (int i = 0; i < 3; ++i) MS_CUDA(cudaStreamSynchronize(h->streams[i]));
        return MELSPEC_OK;
    }
    int64_t per_chunk = std::max<int64_t>(1, target / (ns4 * 4));
    per_chunk = std::min(per_chunk, n_clips);
    if (n_clips >= 3) per_chunk = std::min(per_chunk, (n_clips + 2) / 3);
    int64_t n_chunks = (n_clips + per_chunk - 1) / per_chunk;
    if (h->cfg.frontend == MELSPEC_FRONTEND_KALDI && h->cfg.cmn && n_clips >= h->num_sms && per_chunk < h->num_sms) {
        // Keep every chunk at one clip per SM or more, so that all of them take the kernel with CMN fused in (faster, and the
        // same arithmetic as a device-resident launch of the whole batch: a clip's result does not depend on which CTA or
        // chunk it rides in, so the results are bit-identical).  Bounded: a chunk of num_sms very long clips would make the
        // three staging slots arbitrarily large, so beyond 512 MB of PCM per slot the chunks stay small and CMN runs as the
        // second kernel (same values up to the summation order of the column means).
        const int64_t k = n_clips / h->num_sms;
        const int64_t bytes = ((n_clips + k - 1) / k) * ns4 * 4;
        if (bytes <= ((int64_t)512 << 20)) n_chunks = k;
    }
    // balanced chunks: sizes differ by at most one clip, so no runt chunk falls below the fused-CMN threshold
    const int64_t base = n_clips / n_chunks, rem = n_clips % n_chunks;
    per_chunk = base + (rem ? 1 : 0);
    int32_t rc = ensure_host_resources(h, (size_t)per_chunk * ns4 * 4, (size_t)per_chunk * clip_out * 4,
                                       i16 ? (size_t)per_chunk * ns4 * 2 : 0);
    if (rc) return rc;
    int slot = 0;
    int64_t c0 = 0;
    for (int64_t ci = 0; ci < n_chunks; ++ci, slot = (slot + 1) % 3) {
        const int64_t nc = base + (ci < rem ? 1 : 0);
        cudaStream_t st = h->streams[slot];
        rc = stage_rows(h, slot, h_pcm, i16, c0, nc, clip_stride, 0, n_samples, ns4, st);
        if (rc) return rc;
        rc = launch_device(h, h->d_slot_pcm[slot], nc, ns4, n_samples, F, nullptr, h->d_slot_out[slot], 0, layout, st);
        if (rc) return rc;
        MS_CUDA(cudaMemcpyAsync(h_out + c0 * clip_out, h->d_slot_out[slot], (size_t)nc * clip_out * 4, cudaMemcpyDeviceToHost, st));
        c0 += nc;
    }
    for (int i = 0; i < 3; ++i) MS_CUDA(cudaStreamSynchronize(h->streams[i]));
    return MELSPEC_OK;
}
}  // namespace

int32_t melspec_compute_host(melspec_handle* h, const float* h_pcm, int64_t n_clips, int64_t clip_stride, int64_t n_samples,
                             float* h_out, int32_t layout, int64_t* frames_out) {
    return compute_host_impl(h, h_pcm, false, n_clips, clip_stride, n_samples, h_out, layout, frames_out);
}

int32_t melspec_compute_host_i16(melspec_handle* h, const int16_t* h_pcm, int64_t n_clips, int64_t clip_stride, int64_t n_samples,
                                 float* h_out, int32_t layout, int64_t* frames_out) {
    return compute_host_impl(h, h_pcm, true, n_clips, clip_stride, n_samples, h_out, layout, frames_out);
}

namespace {
// PCM -> Whisper mel -> interleave_frames image -> 8-bit TGA for a batch of clips, pipelined like compute_host_impl: the H2D copy of
// chunk i+1, the kernels of chunk i and the D2H copy of chunk i-1 overlap.  The device -> host side carries one byte per mel value
// instead of four.
int32_t mel_tga_batch_impl(melspec_handle* h, const void* h_pcm, bool i16, int64_t n_clips, int64_t clip_stride, int64_t n_samples,
                           int64_t min_width, uint8_t* h_tga, int64_t tga_stride, int64_t* width_out) {
    if (!h) return fail(MELSPEC_ERR_INVALID_ARG, "handle is null");
    if (h->cfg.frontend != MELSPEC_FRONTEND_WHISPER) return fail(MELSPEC_ERR_UNSUPPORTED, "Whisper frontend only");
    if (n_clips < 0 || n_samples < 0 || clip_stride < 0 || tga_stride < 0) return fail(MELSPEC_ERR_INVALID_ARG, "negative size");
    if (min_width < 0 || (min_width & 1)) return fail(MELSPEC_ERR_INVALID_ARG, "min_width must be even");
    const int64_t F = frames_for(h->cfg, n_samples);
    if (F == 0) return fail(MELSPEC_ERR_INVALID_ARG, "frames is empty");
    const int64_t W = melspec_interleaved_width(F, min_width);
    if (width_out) *width_out = W;
    const int64_t sz = melspec_tga_size(h->cfg.n_mels, W);
    if (sz < 0) return fail(MELSPEC_ERR_INVALID_ARG, "width greater than TARGA max, use chunks (src/quant.rs:17-21)");
    if (n_clips == 0) return MELSPEC_OK;
    if (!h_pcm || !h_tga) return fail(MELSPEC_ERR_INVALID_ARG, "null host pointer");
    if (n_clips > 1 && clip_stride < n_samples) return fail(MELSPEC_ERR_INVALID_ARG, "clip_stride < n_samples");
    if (!tga_stride) tga_stride = sz;
    if (tga_stride < sz) return fail(MELSPEC_ERR_INVALID_ARG, "tga_stride smaller than one image");
    MS_CUDA(cudaSetDevice(h->device));
    const int64_t ns4 = (n_samples + 7) / 8 * 8;
    const int64_t img = (int64_t)h->cfg.n_mels * W;            // floats per interleaved image
    const int64_t dsz = (sz + 15) / 16 * 16;                   // device stride between the TGA images of a chunk
    int64_t per_chunk = std::max<int64_t>(1, ((int64_t)32 << 20) / (ns4 * 4));
    per_chunk = std::min(per_chunk, n_clips);
    if (n_clips >= 3) per_chunk = std::min(per_chunk, (n_clips + 2) / 3);
    per_chunk = std::min<int64_t>(per_chunk, 65535);
    int32_t rc = ensure_host_resources(h, (size_t)per_chunk * ns4 * 4, (size_t)per_chunk * img * 4, i16 ? (size_t)per_chunk * ns4 * 2 : 0);
    if (rc) return rc;
    const size_t tga_need = (size_t)per_chunk * dsz, part_need = sizeof(float2) * (size_t)quantize_partials_per_image(img) * (size_t)per_chunk;
    if (tga_need > h->slot_tga_cap || part_need > h->slot_part_cap) {
        for (int i = 0; i < 3; ++i) {
            if (h->d_slot_tga[i]) cudaFree(h->d_slot_tga[i]);
            if (h->d_slot_part[i]) cudaFree(h->d_slot_part[i]);
            h->d_slot_tga[i] = nullptr; h->d_slot_part[i] = nullptr;
        }
        h->slot_tga_cap = h->slot_part_cap = 0;
        for (int i = 0; i < 3; ++i) {
            MS_CUDA(cudaMalloc(&h->d_slot_tga[i], tga_need));
            MS_CUDA(cudaMalloc(&h->d_slot_part[i], part_need));
        }
        h->slot_tga_cap = tga_need; h->slot_part_cap = part_need;
    }
    int slot = 0;
    for (int64_t c0 = 0; c0 < n_clips; c0 += per_chunk, slot = (slot + 1) % 3) {
        const int64_t nc = std::min(per_chunk, n_clips - c0);
        cudaStream_t st = h->streams[slot];
        rc = stage_rows(h, slot, h_pcm, i16, c0, nc, clip_stride, 0, n_samples, ns4, st);
        if (rc) return rc;
        rc = melspec_compute_interleaved_device(h, h->d_slot_pcm[slot], nc, ns4, n_samples, min_width, h->d_slot_out[slot], 0, st);
        if (rc) return rc;
        rc = quantize_tga_launch(h, h->d_slot_out[slot], nc, img, h->cfg.n_mels, W, h->d_slot_tga[slot], dsz, h->d_slot_part[slot], st);
        if (rc) return rc;
        MS_CUDA(cudaMemcpy2DAsync(h_tga + c0 * tga_stride, (size_t)tga_stride, h->d_slot_tga[slot], (size_t)dsz, (size_t)sz, (size_t)nc,
                                  cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < 3; ++i) MS_CUDA(cudaStreamSynchronize(h->streams[i]));
    return MELSPEC_OK;
}
}  // namespace

int32_t melspec_mel_tga_host_batch(melspec_handle* h, const float* h_pcm, int64_t n_clips, int64_t clip_stride, int64_t n_samples,
                                   int64_t min_width, uint8_t* h_tga, int64_t tga_stride, int64_t* width_out) {
    return mel_tga_batch_impl(h, h_pcm, false, n_clips, clip_stride, n_samples, min_width, h_tga, tga_stride, width_out);
}

int32_t melspec_mel_tga_host_batch_i16(melspec_handle* h, const int16_t* h_pcm, int64_t n_clips, int64_t clip_stride, int64_t n_samples,
                                       int64_t min_width, uint8_t* h_tga, int64_t tga_stride, int64_t* width_out) {
    return mel_tga_batch_impl(h, h_pcm, true, n_clips, clip_stride, n_samples, min_width, h_tga, tga_stride, width_out);
}

int32_t melspec_convert_i16_device(melspec_handle* h, const int16_t* d_in, int64_t n_rows, int64_t in_stride, int64_t n_samples,
                                   float* d_out, int64_t out_stride, void* stream) {
    if (!h) return fail(MELSPEC_ERR_INVALID_ARG, "handle is null");
    if (n_rows < 0 || in_stride < 0 || out_stride < 0 || n_samples < 0 || n_samples > 0x7fffffff) return fail(MELSPEC_ERR_INVALID_ARG, "bad size");
    if (n_rows == 0 || n_samples == 0) return MELSPEC_OK;
    if (!d_in || !d_out) return fail(MELSPEC_ERR_INVALID_ARG, "null device pointer");
    if (n_rows > 1 && (in_stride < n_samples || out_stride < n_samples)) return fail(MELSPEC_ERR_INVALID_ARG, "stride < n_samples");
    MS_CUDA(cudaSetDevice(h->device));
    return launch_convert_i16(h, d_in, n_rows, in_stride ? in_stride : n_samples, n_samples, d_out, out_stride ? out_stride : n_samples,
                              (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------ streaming
int32_t melspec_stream_create(melspec_handle* h, int64_t max_chunk_samples, melspec_stream** out) {
    if (!h || !out) return fail(MELSPEC_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if (max_chunk_samples <= 0) return fail(MELSPEC_ERR_INVALID_ARG, "max_chunk_samples must be positive");
    // Spectrogram::add / RingBuffer are the Whisper STFT (src/stft.rs:48-86, src/rb.rs:86-121).  A Kaldi stream would need the
    // CMN and the pre-emphasis look-back carried across pushes, a NeMo stream the centre padding: neither exists in the
    // reference.  hop > fft: the reference's overlap buffer arithmetic panics (stft.rs:50-55), here it is an error.
    if (h->cfg.frontend != MELSPEC_FRONTEND_WHISPER)
        return fail(MELSPEC_ERR_UNSUPPORTED, "streaming is defined for the Whisper frontend only (src/stft.rs:48-86)");
    if (h->cfg.hop > h->cfg.fft) return fail(MELSPEC_ERR_UNSUPPORTED, "streaming needs hop_size <= fft_size (src/stft.rs:48-59)");
    MS_CUDA(cudaSetDevice(h->device));
    melspec_stream* s = new (std::nothrow) melspec_stream();
    if (!s) return fail(MELSPEC_ERR_CUDA, "out of host memory");
    s->h = h;
    s->max_chunk = max_chunk_samples;
    const Resolved& c = h->cfg;
    // a push is cut into pieces of <= 4 s of 16 kHz audio; piece k+1's H2D copy overlaps piece k's kernel and D2H copy
    // (pieces of up to 2^22 samples = 4.4 min of 16 kHz audio: far below that a piece's kernel + copies are launch-latency bound, ~25 us each)
    static const int64_t piece_cap = [] {
        const char* e = std::getenv("MELSPEC_STREAM_PIECE");
        const int64_t v = e ? std::atoll(e) : 0;
        return v >= 1024 ? v : (int64_t)(1 << 22);
    }();
    s->piece = std::min<int64_t>(max_chunk_samples, piece_cap);
    s->cap = (s->piece + c.frame_len + c.hop + 8 + 3) / 4 * 4;
    s->to_skip = (int64_t)((c.frame_len + c.hop - 1) / c.hop) * c.hop - c.frame_len;   // c = ceil(N/H)*H - N
    s->out_cap_frames = s->cap / c.hop + 2;
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaMalloc(&s->d_buf[i], (size_t)s->cap * 4);
        if (e == cudaSuccess) e = cudaMemset(s->d_buf[i], 0, (size_t)s->cap * 4);
        if (e == cudaSuccess) e = cudaMalloc(&s->d_out[i], (size_t)s->out_cap_frames * c.n_mels * 4);
        if (e == cudaSuccess) e = cudaHostAlloc(&s->h_pin_in[i], (size_t)s->piece * 4, cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaHostAlloc(&s->h_pin_out[i], (size_t)s->out_cap_frames * c.n_mels * 4, cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_h2d[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_free[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_d2h[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->compute_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        melspec_stream_destroy(s);
        return fail_cuda(e, "melspec_stream_create");
    }
    *out = s;
    return MELSPEC_OK;
}

int32_t melspec_stream_reset(melspec_stream* s) {
    if (!s) return fail(MELSPEC_ERR_INVALID_ARG, "stream is null");
    const Resolved& c = s->h->cfg;
    cudaSetDevice(s->h->device);
    cudaStreamSynchronize(s->copy_stream);
    cudaStreamSynchronize(s->compute_stream);
    s->buffered = 0;
    s->cur = 0;
    s->idx = 0;
    s->used_free[0] = s->used_free[1] = false;
    s->to_skip = (int64_t)((c.frame_len + c.hop - 1) / c.hop) * c.hop - c.frame_len;
    return MELSPEC_OK;
}

int32_t melspec_stream_push(melspec_stream* s, const float* h_samples, int64_t n, float* h_out, int64_t out_capacity_frames,
                            int64_t* frames_emitted) {
    if (frames_emitted) *frames_emitted = 0;
    if (!s) return fail(MELSPEC_ERR_INVALID_ARG, "stream is null");
    if (n < 0 || n > s->max_chunk) return fail(MELSPEC_ERR_INVALID_ARG, "chunk larger than max_chunk_samples");
    if (n == 0) return MELSPEC_OK;
    if (!h_samples) return fail(MELSPEC_ERR_INVALID_ARG, "null samples");
    melspec_handle* h = s->h;
    const Resolved& c = h->cfg;
    MS_CUDA(cudaSetDevice(h->device));
    // the first c samples of the stream never reach a frame (src/stft.rs:61-66 fed whole hops)
    const int64_t skip = std::min(s->to_skip, n);
    s->to_skip -= skip;
    h_samples += skip;
    n -= skip;
    if (n == 0) return MELSPEC_OK;
    // frames are emitted on whole-hop boundaries: frame k ends at stream offset k*hop + N past the carried tail
    const int64_t total = s->buffered + n;
    const int64_t nf_total = total >= c.frame_len ? (total - c.frame_len) / c.hop + 1 : 0;
    if (nf_total > out_capacity_frames) return fail(MELSPEC_ERR_INVALID_ARG, "out_capacity_frames too small for this push");
    if (nf_total > 0 && !h_out) return fail(MELSPEC_ERR_INVALID_ARG, "null output");
    // pinned (or registered) caller buffers are DMA'd directly; pageable ones go through the stream's own pinned slots
    auto is_pinned = [](const void* p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return a.type == cudaMemoryTypeHost;
    };
    const bool pin_in = is_pinned(h_samples), pin_out = nf_total > 0 && is_pinned(h_out);
    struct Pending { int slot; int64_t frames, at; };
    Pending pend[2] = {{-1, 0, 0}, {-1, 0, 0}};
    int64_t done_frames = 0;
    int k = 0;
    for (int64_t off = 0; off < n; off += s->piece, ++k) {
        const int64_t m = std::min(s->piece, n - off);
        const int slot = k & 1;
        float* buf = s->d_buf[s->cur];
        // the pageable path reuses staging slot `slot`: its previous H2D (two pieces ago) must be done, and the output slot
        // must have been copied out to the caller
        const float* src = h_samples + off;
        if (!pin_in) {
            if (k >= 2) MS_CUDA(cudaEventSynchronize(s->ev_h2d[slot]));
            std::memcpy(s->h_pin_in[slot], h_samples + off, (size_t)m * 4);
            src = s->h_pin_in[slot];
        }
        if (!pin_out && pend[slot].slot >= 0) {
            MS_CUDA(cudaEventSynchronize(s->ev_d2h[slot]));
            std::memcpy(h_out + pend[slot].at * c.n_mels, s->h_pin_out[slot], (size_t)pend[slot].frames * c.n_mels * 4);
            pend[slot].slot = -1;
        }
        // H2D of this piece behind the carried tail of the current device buffer (side stream); the buffer was last read
        // by the kernel two pieces ago
        if (s->used_free[s->cur]) MS_CUDA(cudaStreamWaitEvent(s->copy_stream, s->ev_free[s->cur], 0));
        MS_CUDA(cudaMemcpyAsync(buf + s->buffered, src, (size_t)m * 4, cudaMemcpyHostToDevice, s->copy_stream));
        MS_CUDA(cudaEventRecord(s->ev_h2d[slot], s->copy_stream));
        MS_CUDA(cudaStreamWaitEvent(s->compute_stream, s->ev_h2d[slot], 0));
        const int64_t have = s->buffered + m;
        const int64_t nf = have >= c.frame_len ? (have - c.frame_len) / c.hop + 1 : 0;
        if (nf > 0) {
            const int64_t ns4 = std::min<int64_t>((have + 3) / 4 * 4, s->cap);
            int32_t rc = launch_device(h, buf, 1, s->cap, ns4, nf, nullptr, s->d_out[slot], 0, MELSPEC_LAYOUT_FRAME_MAJOR, s->compute_stream);
            if (rc) return rc;
            float* dst = pin_out ? h_out + done_frames * c.n_mels : s->h_pin_out[slot];
            MS_CUDA(cudaMemcpyAsync(dst, s->d_out[slot], (size_t)nf * c.n_mels * 4, cudaMemcpyDeviceToHost, s->compute_stream));
            MS_CUDA(cudaEventRecord(s->ev_d2h[slot], s->compute_stream));
            if (!pin_out) pend[slot] = {slot, nf, done_frames};
            // carry the tail (everything from the start of the next frame) into the other buffer
            const int64_t consumed = nf * c.hop, keep = have - consumed;
            const int other = s->cur ^ 1;
            if (s->used_free[other]) MS_CUDA(cudaStreamWaitEvent(s->compute_stream, s->ev_free[other], 0));
            MS_CUDA(cudaMemcpyAsync(s->d_buf[other], buf + consumed, (size_t)keep * 4, cudaMemcpyDeviceToDevice, s->compute_stream));
            MS_CUDA(cudaEventRecord(s->ev_free[s->cur], s->compute_stream));
            s->used_free[s->cur] = true;
            s->cur = other;
            s->buffered = keep;
            done_frames += nf;
        } else {
            s->buffered = have;
        }
    }
    MS_CUDA(cudaStreamSynchronize(s->compute_stream));
    MS_CUDA(cudaStreamSynchronize(s->copy_stream));
    for (int i = 0; i < 2; ++i)
        if (pend[i].slot >= 0)
            std::memcpy(h_out + pend[i].at * c.n_mels, s->h_pin_out[i], (size_t)pend[i].frames * c.n_mels * 4);
    if (frames_emitted) *frames_emitted = done_frames;
    return MELSPEC_OK;
}

// Spectrogram::add (src/stft.rs:48-86): the hop buffer advances by one whole hop per call whatever the chunk length; fed through
// the whole-hop machinery above, frame j of that machinery is exactly the reference's frame of call j (both cover the last
// fft_size samples of the zero-padded hop sequence), and the reference hands it out once idx >= fft_size.
int32_t melspec_stream_push_hop(melspec_stream* s, const float* h_samples, int64_t n, float* h_out_frame, int32_t* emitted) {
    if (emitted) *emitted = 0;
    if (!s) return fail(MELSPEC_ERR_INVALID_ARG, "stream is null");
    const Resolved& c = s->h->cfg;
    if (n < 0 || n > c.hop) return fail(MELSPEC_ERR_INVALID_ARG, "frames must be <= hop_size");   // assert at src/stft.rs:53
    if (n > 0 && !h_samples) return fail(MELSPEC_ERR_INVALID_ARG, "null samples");
    if (!h_out_frame) return fail(MELSPEC_ERR_INVALID_ARG, "null output");
    if (s->max_chunk < c.hop) return fail(MELSPEC_ERR_INVALID_ARG, "stream was created with max_chunk_samples < hop_size");
    try {
        s->hopbuf.assign((size_t)c.hop, 0.f);                                                     // zero pad, src/stft.rs:56-59
    } catch (...) {
        return fail(MELSPEC_ERR_CUDA, "out of host memory");
    }
    if (n > 0) std::memcpy(s->hopbuf.data(), h_samples, (size_t)n * 4);
    int64_t got = 0;
    int32_t rc = melspec_stream_push(s, s->hopbuf.data(), c.hop, h_out_frame, 1, &got);
    if (rc) return rc;
    s->idx += (uint64_t)n;                                                                        // wrapping_add, src/stft.rs:64
    if (emitted) *emitted = (got == 1 && s->idx >= (uint64_t)c.fft) ? 1 : 0;                      // src/stft.rs:66
    return MELSPEC_OK;
}

void melspec_stream_destroy(melspec_stream* s) {
    if (!s) return;
    cudaSetDevice(s->h->device);
    if (s->copy_stream) cudaStreamSynchronize(s->copy_stream);
    if (s->compute_stream) cudaStreamSynchronize(s->compute_stream);
    for (int i = 0; i < 2; ++i) {
        if (s->d_buf[i]) cudaFree(s->d_buf[i]);
        if (s->d_out[i]) cudaFree(s->d_out[i]);
        if (s->h_pin_in[i]) cudaFreeHost(s->h_pin_in[i]);
        if (s->h_pin_out[i]) cudaFreeHost(s->h_pin_out[i]);
        if (s->ev_h2d[i]) cudaEventDestroy(s->ev_h2d[i]);
        if (s->ev_free[i]) cudaEventDestroy(s->ev_free[i]);
        if (s->ev_d2h[i]) cudaEventDestroy(s->ev_d2h[i]);
    }
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    if (s->compute_stream) cudaStreamDestroy(s->compute_stream);
    delete s;
}

}  // extern "C"
