// melspec_generic.cuh — the general plan: any fft_size / hop / frame length / n_mels <= 128 for the three frontends.
//
// The reference accepts arbitrary sizes (`Spectrogram::compute_mel_spectrogram_cpu(samples, fft_size, hop_size, sr, n_mels)`,
// src/stft.rs:119-138; `FbankConfig` frame lengths / sample rates, src/fbank.rs:25-82; `BatchLogMelConfig`,
// src/mel.rs:171-214) and plans its FFT with rustfft, which takes any length.  The two specialised kernels of
// melspec_kernels.cuh cover the configurations the reference's own tests, goldens and BASELINE.json name (fft 400 and
// fft 512 / 400-sample frames at hop 160); every other configuration runs here, so that the C ABI is a drop-in for the
// whole parameter space of the path and not only for its headline points.  Same fusion (PCM read once, mel written once,
// nothing in between touches HBM), simpler schedule:
//
//   * one warp per frame.  Even N: the frame's even / odd samples are the real / imaginary part of one complex N/2-point
//     sequence z[n] = x[2n] + i x[2n+1]; X[k] = E[k] + W_N^k O[k] with E = (Z[k] + conj Z[N/2-k])/2,
//     O = (Z[k] - conj Z[N/2-k])/2i.  Odd N: a complex N-point transform of the real frame.  (Unlike the two-frames-per-
//     transform packing of the specialised kernels this keeps every frame's rounding noise relative to its *own* level,
//     whatever its neighbours contain.)
//   * the frontend prologue while the samples are on their way from global memory: periodic Hann (Whisper,
//     src/stft.rs:141-169), DC removal + pre-emphasis with look-back + Povey window + zero padding (Kaldi,
//     src/fbank.rs:166-190), whole-waveform pre-emphasis + centre padding + centred symmetric Hann (NeMo,
//     src/mel.rs:685-719);
//   * a mixed-radix Stockham autosort FFT in the warp's private shared-memory ping-pong buffers (radix 4 and 2 in
//     registers, any other prime factor r as an r-term sum per output, balanced over (butterfly, output) pairs so that a
//     large prime factor — even a prime N — is slow but correct); twiddles W_N^k come from one f64-built table per CTA;
//   * power (or magnitude) of bins 0..N/2, banded projection in ascending bin order from a CSR table (one lane per mel
//     row, the reference's own summation order, src/mel.rs:106-168), log / floor / guard and the Whisper per-frame clamp
//     (src/mel.rs:645-654) with a warp max, stores in the caller's layout.
//
// Only `__syncwarp()` inside the loop.  HBM traffic is the algorithmic 4*hop + 4*n_mels bytes per frame (overlapping
// frames of neighbouring warps are served by L1/L2).
#pragma once
#include "melspec_kernels.cuh"

namespace melspec {

constexpr int kMaxStages = 24;

struct GParams {
    const float2* tw;       // [N] W_N^k = (cos(2 pi k/N), -sin(2 pi k/N)), rounded once from f64
    const float* window;    // [frame_len]
    const int* bands;       // [n_mels][3]: first bin, number of bins, offset into weights
    const float* weights;   // band weights, f32 of the f64 filterbank
    int N;                  // frame transform length (fft_size)
    int Nf;                 // length of the complex transform actually run: N/2 (even N) or N (odd N)
    int n_stages;           // radix schedule of Nf
    int mode;               // 0 Whisper, 1 Kaldi, 2 NeMo
    int use_power, use_log;
    long long n_units;      // frames_per_clip * n_clips
    int radix[kMaxStages];
    int sshift[kMaxStages]; // log2 of the stage's stride s when it is a power of two (shift instead of an integer division), else -1
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, fmaf(a.x, b.y, a.y * b.x)); }

__global__ void __launch_bounds__(256) melspec_generic_kernel(const KParams p, const GParams g) {
    extern __shared__ __align__(16) unsigned char gsm[];
    const int N = g.N, Nf = g.Nf, tmul = N / Nf;   // W_Nf^k = W_N^(tmul k)
    const bool packed = Nf != N;
    const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2* s_tw = reinterpret_cast<float2*>(gsm);
    float2* buf0 = s_tw + N + (size_t)warp * 2 * Nf;
    float2* buf1 = buf0 + Nf;
    for (int i = threadIdx.x; i < N; i += blockDim.x) s_tw[i] = g.tw[i];
    __syncthreads();

    const int L = p.frame_len;
    const float inv_len = 1.0f / (float)L;
    const int nb = N / 2;   // last bin formed

    for (long long u = (long long)blockIdx.x * nw + warp; u < g.n_units; u += (long long)gridDim.x * nw) {
        const int clip = (int)(u / p.frames_per_clip);
        const int f0 = (int)(u - (long long)clip * p.frames_per_clip);
        int len = p.n_samples;
        if (g.mode != 2 && p.lens) {
            len = min(p.lens[clip], p.n_samples);
            const int nfr = len < L ? 0 : (len - L) / p.hop + 1;
            if (f0 >= nfr) continue;   // frames past a short clip's own frame count are left untouched
        }
        const float* x = p.pcm + (long long)clip * p.clip_stride;
        const long long sa = (long long)f0 * p.hop + p.frame_offset;
        auto at = [&](long long i) -> float { return (i >= 0 && i < len) ? __ldg(x + i) : 0.f; };

        // ------------------------------------------------------------------ prologue: windowed sample n of the frame
        float mu = 0.f;
        if (g.mode == 1) {   // per-frame mean (src/fbank.rs:166-170)
            float s0 = 0.f;
            for (int n = lane; n < L; n += 32) s0 += at(sa + n);
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            mu = s0 * inv_len;
        }
        auto sample = [&](int n) -> float {
            if (n >= L) return 0.f;   // zero padding up to the FFT size
            const float w = __ldg(g.window + n);
            if (g.mode == 0) return at(sa + n) * w;
            if (g.mode == 1) {        // src/fbank.rs:172-190: frame 0 of a clip has no look-back sample
                float a = at(sa + n) - mu;
                if (n > 0 || sa > 0) a = fmaf(-p.preemph, at(sa + n - 1) - mu, a);
                return a * w;
            }
            const long long ia = sa + n;   // src/mel.rs:696-706: wave[i] = x[i] - c x[i-1] (i >= 1), zero outside the clip
            return (ia >= 0 && ia < len) ? fmaf(-p.preemph, at(ia - 1), at(ia)) * w : 0.f;
        };
        if (packed) for (int n = lane; n < Nf; n += 32) buf0[n] = make_float2(sample(2 * n), sample(2 * n + 1));
        else        for (int n = lane; n < Nf; n += 32) buf0[n] = make_float2(sample(n), 0.f);
        __syncwarp();

        // ------------------------------------------------------------------ Stockham autosort FFT (decimation in frequency)
        // stage with radix r on sub-transforms of length n = r m, stride s (product of the earlier radices):
        //   y[q + s (r p + j)] = W_Nf^(p j s) * sum_i x[q + s (p + m i)] W_r^(i j),   p < m, q < s, j < r
        float2* src = buf0;
        float2* dst = buf1;
        int ncur = Nf, s = 1;
        for (int st = 0; st < g.n_stages; ++st) {
            const int r = g.radix[st], m = ncur / r, sh = g.sshift[st];
            if (r == 4) {
                for (int bfly = lane; bfly < Nf / 4; bfly += 32) {
                    const int pp = sh >= 0 ? bfly >> sh : bfly / s, q = bfly - pp * s;
                    const float2* xi = src + q + s * pp;
                    const float2 a0 = xi[0], a1 = xi[(size_t)s * m], a2 = xi[(size_t)2 * s * m], a3 = xi[(size_t)3 * s * m];
                    const float2 t0 = make_float2(a0.x + a2.x, a0.y + a2.y), t1 = make_float2(a0.x - a2.x, a0.y - a2.y);
                    const float2 t2 = make_float2(a1.x + a3.x, a1.y + a3.y), t3 = make_float2(a1.x - a3.x, a1.y - a3.y);
                    float2* yo = dst + q + (size_t)s * 4 * pp;
                    const int tws = pp * s * tmul;
                    yo[0] = make_float2(t0.x + t2.x, t0.y + t2.y);
                    yo[s] = cmul(make_float2(t1.x + t3.y, t1.y - t3.x), s_tw[tws]);          // a0 - i a1 - a2 + i a3
                    yo[2 * s] = cmul(make_float2(t0.x - t2.x, t0.y - t2.y), s_tw[2 * tws]);
                    yo[3 * s] = cmul(make_float2(t1.x - t3.y, t1.y + t3.x), s_tw[3 * tws]);  // a0 + i a1 - a2 - i a3
                }
            } else if (r == 2) {
                for (int bfly = lane; bfly < Nf / 2; bfly += 32) {
                    const int pp = sh >= 0 ? bfly >> sh : bfly / s, q = bfly - pp * s;
                    const float2 a0 = src[q + s * pp], a1 = src[q + s * (pp + m)];
                    float2* yo = dst + q + (size_t)s * 2 * pp;
                    yo[0] = make_float2(a0.x + a1.x, a0.y + a1.y);
                    yo[s] = cmul(make_float2(a0.x - a1.x, a0.y - a1.y), s_tw[pp * s * tmul]);
                }
            } else {   // any other prime factor: one (butterfly, output) pair per work item
                const int wr = N / r;   // W_r^e = W_N^(e N/r)
                for (int e = lane; e < Nf; e += 32) {
                    const int bfly = e / r, j = e - bfly * r;
                    const int pp = sh >= 0 ? bfly >> sh : bfly / s, q = bfly - pp * s;
                    const float2* xi = src + q + s * pp;
                    float2 acc = xi[0];
                    int idx = 0;
                    for (int i = 1; i < r; ++i) {
                        idx += j;
                        if (idx >= r) idx -= r;
                        const float2 v = xi[(size_t)s * m * i], w = s_tw[idx * wr];
                        acc.x = fmaf(v.x, w.x, fmaf(-v.y, w.y, acc.x));
                        acc.y = fmaf(v.x, w.y, fmaf(v.y, w.x, acc.y));
                    }
                    dst[q + (size_t)s * ((size_t)r * pp + j)] = cmul(acc, s_tw[pp * j * s * tmul]);
                }
            }
            __syncwarp();
            float2* t = src; src = dst; dst = t;
            ncur = m; s *= r;
        }

        // ------------------------------------------------------------------ power (or magnitude) of bins 0..N/2 -> pw[k]
        float* pw = reinterpret_cast<float*>(dst);
        for (int k = lane; k <= nb; k += 32) {
            float xr, xi;
            if (packed) {   // X[k] = E[k] + W_N^k O[k]
                const float2 zk = src[k == Nf ? 0 : k], zm = src[(k == 0 || k == Nf) ? 0 : Nf - k];
                const float er = 0.5f * (zk.x + zm.x), ei = 0.5f * (zk.y - zm.y);
                const float orr = 0.5f * (zk.y + zm.y), oi = -0.5f * (zk.x - zm.x);
                const float2 w = k == Nf ? make_float2(-1.f, 0.f) : s_tw[k];
                xr = er + (orr * w.x - oi * w.y);
                xi = ei + fmaf(orr, w.y, oi * w.x);
            } else {
                xr = src[k].x; xi = src[k].y;
            }
            float e = fmaf(xr, xr, xi * xi);
            if (!g.use_power) e = sqrtf(e);   // src/fbank.rs:197-203
            pw[k] = e;
        }
        __syncwarp();

        // ------------------------------------------------------------------ banded projection + log + stores
        float v[kMaxMpl];
        float mx = -INFINITY;
#pragma unroll
        for (int sl = 0; sl < kMaxMpl; ++sl) {
            const int mrow = lane + 32 * sl;
            v[sl] = -INFINITY;
            if (mrow < p.n_mels) {
                const int b0 = __ldg(g.bands + 3 * mrow), cnt = __ldg(g.bands + 3 * mrow + 1), wo = __ldg(g.bands + 3 * mrow + 2);
                float e = 0.f;
                for (int i = 0; i < cnt; ++i) e = fmaf(__ldg(g.weights + wo + i), pw[b0 + i], e);
                if (g.mode == 2) e = logf(e + p.log_add);                    // ln(E + guard), src/mel.rs:365-368
                else if (g.mode == 1) {                                      // max(E, floor), optional ln, src/fbank.rs:207-221
                    e = fmaxf(e, p.floor_val);
                    if (g.use_log) e = logf(e);
                } else e = log10f(fmaxf(e, p.floor_val));                    // log10(max(E, 1e-10)), src/mel.rs:148-168
                v[sl] = e;
                mx = fmaxf(mx, e);
            }
        }
        if (p.normalize) {   // per-frame clamp to max - 8, then (x + 4)/4 (src/mel.rs:645-654)
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        float* oc = p.out + (long long)clip * p.out_clip_stride;
#pragma unroll
        for (int sl = 0; sl < kMaxMpl; ++sl) {
            const int mrow = lane + 32 * sl;
            if (mrow < p.n_mels) {
                float y = v[sl];
                if (p.normalize) y = (fmaxf(y, mx - 8.0f) + 4.0f) * 0.25f;
                if (p.layout == 0) oc[(long long)f0 * p.n_mels + mrow] = y;
                else oc[(long long)mrow * p.out_row_stride + f0] = y;
            }
        }
        __syncwarp();   // the next unit overwrites both buffers
    }
}

}  // namespace melspec
