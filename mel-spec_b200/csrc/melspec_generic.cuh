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
//   * a mixed-radix Stockham autosort FFT in the warp's private shared-memory ping-pong buffers: radix 8, 4, 2, 3 and 5
//     butterflies in registers, any other prime factor r as an r-term sum per output, balanced over (butterfly, output)
//     pairs so that a large prime factor — even a prime N — is slow but correct.  Element i lives at i + (i >> 4): with
//     that one-in-sixteen padding the strided 64-bit stores of the early stages spread over all 16 bank pairs of a
//     half-warp (modelled: 224 instead of 512 wavefronts per 512-point transform together with radix 8).  Twiddles
//     W_N^k come from one f64-built table per CTA;
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
    int vec2;               // Whisper prologue may use 64-bit loads (8-byte aligned rows, even hop and frame offset)
    int center;             // NeMo: frames are centred (frame count len/hop + 1), for per-clip lengths
    // pair form (melspec_generic2.cuh): the same weights transposed and zero padded, [slot][entry][lane] with kmax[slot] entries per
    // slot (slot = mel row / 32, lane = mel row % 32), staged in shared memory: conflict-free reads, warp-uniform trip counts
    const float* weights_t;
    const int* starts_t;    // [n_mels] first power row of the mel row's window: at or up to 15 rows before its band, chosen so that the
                            // 16 lanes of a half-warp start at 16 different rows mod 16 (conflict-free 64-bit reads of the power pairs)
    int n_weights_t;        // 32 * (kmax[0] + .. + kmax[3]); 0: table too large for shared memory, the pair form is not used
    int kmax[4];
};

// physical position of element i in a ping-pong buffer (one pad slot per 16 elements)
__device__ __forceinline__ int gph(int i) { return i + (i >> 4); }
__host__ __device__ inline int generic_buf_elems(int nf) { return ((nf + (nf >> 4) + 2) + 1) & ~1; }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, fmaf(a.x, b.y, a.y * b.x)); }
// complex add / subtract as one packed FADD2 (sm_100): half the issue slots of two scalar FADDs
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return add2(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return sub2(a, b); }

// forward 4-point DFT, natural order out
__device__ __forceinline__ void gdft4(float2 b0, float2 b1, float2 b2, float2 b3, float2& y0, float2& y1, float2& y2, float2& y3) {
    const float2 s02 = cadd(b0, b2), d02 = csub(b0, b2), s13 = cadd(b1, b3), d13 = csub(b1, b3);
    y0 = cadd(s02, s13);
    y1 = make_float2(d02.x + d13.y, d02.y - d13.x);   // d02 - i d13
    y2 = csub(s02, s13);
    y3 = make_float2(d02.x - d13.y, d02.y + d13.x);   // d02 + i d13
}

// forward 8-point DFT in place, natural order (one radix-2 layer with W_8^k, then two 4-point DFTs)
__device__ __forceinline__ void gdft8(float2 (&a)[8]) {
    constexpr float c = 0.70710678118654752f;
    const float2 t0 = cadd(a[0], a[4]), t4 = csub(a[0], a[4]);
    const float2 t1 = cadd(a[1], a[5]), d1 = csub(a[1], a[5]);
    const float2 t2 = cadd(a[2], a[6]), d2 = csub(a[2], a[6]);
    const float2 t3 = cadd(a[3], a[7]), d3 = csub(a[3], a[7]);
    const float2 t5 = make_float2(c * (d1.x + d1.y), c * (d1.y - d1.x));    // d1 W_8
    const float2 t6 = make_float2(d2.y, -d2.x);                             // d2 W_8^2 = -i d2
    const float2 t7 = make_float2(c * (d3.y - d3.x), -c * (d3.x + d3.y));   // d3 W_8^3
    gdft4(t0, t1, t2, t3, a[0], a[2], a[4], a[6]);
    gdft4(t4, t5, t6, t7, a[1], a[3], a[5], a[7]);
}

// forward R-point DFT for R = 3, 5 from compile-time roots of unity (r-term sums in registers)
// (cos, sin)(2 pi e / R) as compile-time constants (the loops below are fully unrolled, so `e` is a constant)
template <int R> __host__ __device__ constexpr float groot_cos(int e) {
    return R == 3 ? (e == 0 ? 1.f : -0.5f)
                  : (e == 0 ? 1.f : (e == 1 || e == 4) ? 0.30901699437494742f : -0.80901699437494742f);
}
template <int R> __host__ __device__ constexpr float groot_sin(int e) {
    return R == 3 ? (e == 0 ? 0.f : e == 1 ? 0.86602540378443865f : -0.86602540378443865f)
                  : (e == 0 ? 0.f : e == 1 ? 0.95105651629515357f : e == 2 ? 0.58778525229247313f
                                   : e == 3 ? -0.58778525229247313f : -0.95105651629515357f);
}
template <int R>
__device__ __forceinline__ void gdft_small(const float2 (&a)[R], float2 (&y)[R]) {
#pragma unroll
    for (int j = 0; j < R; ++j) {
        float2 acc = a[0];
#pragma unroll
        for (int i = 1; i < R; ++i) {
            const int e = (i * j) % R;                       // W_R^e = (cos, -sin)(2 pi e / R)
            const float wc = groot_cos<R>(e), ws = -groot_sin<R>(e);
            acc.x = fmaf(a[i].x, wc, fmaf(-a[i].y, ws, acc.x));
            acc.y = fmaf(a[i].x, ws, fmaf(a[i].y, wc, acc.y));
        }
        y[j] = acc;
    }
}

__global__ void __launch_bounds__(512, 2) melspec_generic_kernel(const KParams p, const GParams g) {
    extern __shared__ __align__(16) unsigned char gsm[];
    const int N = g.N, Nf = g.Nf, tmul = N / Nf;   // W_Nf^k = W_N^(tmul k)
    const bool packed = Nf != N;
    const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2* s_tw = reinterpret_cast<float2*>(gsm);
    const int nfp = generic_buf_elems(Nf);
    // [twiddles | transposed band weights (round 2, when they fit: n_weights_t > 0) | per-warp ping-pong buffers]
    float* s_wts = reinterpret_cast<float*>(s_tw + N);
    const int wt_floats = (g.n_weights_t + 1) & ~1;
    float2* buf0 = s_tw + N + wt_floats / 2 + (size_t)warp * 2 * nfp;
    float2* buf1 = buf0 + nfp;
    for (int i = threadIdx.x; i < N; i += blockDim.x) s_tw[i] = g.tw[i];
    for (int i = threadIdx.x; i < g.n_weights_t; i += blockDim.x) s_wts[i] = g.weights_t[i];
    __syncthreads();
    const bool wsm = g.n_weights_t > 0;   // projection from the shared-memory table (warp-uniform trip counts, conflict-free reads)
    int kpad = 0;
    for (int sl = 0; sl < kMaxMpl; ++sl) kpad = max(kpad, g.kmax[sl]);

    const int L = p.frame_len;
    const float inv_len = 1.0f / (float)L;
    const int nb = N / 2;   // last bin formed

    for (long long u = (long long)blockIdx.x * nw + warp; u < g.n_units; u += (long long)gridDim.x * nw) {
        const int clip = (int)(u / p.frames_per_clip);
        const int f0 = (int)(u - (long long)clip * p.frames_per_clip);
        int len = p.n_samples;
        if (p.lens) {
            len = min(p.lens[clip], p.n_samples);
            if (g.mode != 2) {
                const int nfr = len < L ? 0 : (len - L) / p.hop + 1;
                if (f0 >= nfr) continue;   // frames past a short clip's own frame count are left untouched
            } else {
                // NeMo, ragged batch: the clip's own frame count (src/mel.rs:387-395); the columns past it are zeros, like the
                // pad_to columns of the reference's feature matrix (src/mel.rs:336)
                const int nfr = len <= 0 ? 0 : g.center ? len / p.hop + 1 : (len < N ? 0 : (len - N) / p.hop + 1);
                if (f0 >= nfr) {
                    float* oc0 = p.out + (long long)clip * p.out_clip_stride;
                    for (int mrow = lane; mrow < p.n_mels; mrow += 32) {
                        if (p.layout == 0) oc0[(long long)f0 * p.n_mels + mrow] = 0.f;
                        else oc0[(long long)mrow * p.out_row_stride + f0] = 0.f;
                    }
                    continue;
                }
            }
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
        if (packed && g.mode == 0 && g.vec2 && sa + N <= len) {   // whole frame inside the clip, 8-byte aligned pairs
            const float2* x2 = reinterpret_cast<const float2*>(x + sa);
            const float2* w2 = reinterpret_cast<const float2*>(g.window);
            for (int n = lane; n < Nf; n += 32) {
                const float2 v = __ldg(x2 + n), w = __ldg(w2 + n);
                buf0[gph(n)] = make_float2(v.x * w.x, v.y * w.y);
            }
        } else if (packed) {
            for (int n = lane; n < Nf; n += 32) buf0[gph(n)] = make_float2(sample(2 * n), sample(2 * n + 1));
        } else {
            for (int n = lane; n < Nf; n += 32) buf0[gph(n)] = make_float2(sample(n), 0.f);
        }
        __syncwarp();

        // ------------------------------------------------------------------ Stockham autosort FFT (decimation in frequency)
        // stage with radix r on sub-transforms of length n = r m, stride s (product of the earlier radices):
        //   y[q + s (r p + j)] = W_Nf^(p j s) * sum_i x[q + s (p + m i)] W_r^(i j),   p < m, q < s, j < r
        float2* src = buf0;
        float2* dst = buf1;
        int ncur = Nf, s = 1;
        for (int st = 0; st < g.n_stages; ++st) {
            const int r = g.radix[st], m = ncur / r, sh = g.sshift[st];
            const int sm = s * m;   // input stride between the r legs of a butterfly
            if (r == 8) {
                for (int bfly = lane; bfly < Nf / 8; bfly += 32) {
                    const int pp = sh >= 0 ? bfly >> sh : bfly / s, q = bfly - pp * s;
                    const int bi = q + s * pp, bo = q + s * 8 * pp;
                    float2 a[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) a[i] = src[gph(bi + sm * i)];
                    gdft8(a);
                    if (m > 1) {   // W^(j p s), j = 1..7, from three table reads
                        const int tws = pp * s * tmul;
                        const float2 w1 = s_tw[tws], w2 = s_tw[2 * tws], w4 = s_tw[4 * tws];
                        const float2 w3 = cmul(w1, w2), w5 = cmul(w4, w1), w6 = cmul(w4, w2);
                        a[1] = cmul(a[1], w1); a[2] = cmul(a[2], w2); a[3] = cmul(a[3], w3); a[4] = cmul(a[4], w4);
                        a[5] = cmul(a[5], w5); a[6] = cmul(a[6], w6); a[7] = cmul(a[7], cmul(w4, w3));
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) dst[gph(bo + s * j)] = a[j];
                }
            } else if (r == 4) {
                for (int bfly = lane; bfly < Nf / 4; bfly += 32) {
                    const int pp = sh >= 0 ? bfly >> sh : bfly / s, q = bfly - pp * s;
                    const int bi = q + s * pp, bo = q + s * 4 * pp;
                    float2 y0, y1, y2, y3;
                    gdft4(src[gph(bi)], src[gph(bi + sm)], src[gph(bi + 2 * sm)], src[gph(bi + 3 * sm)], y0, y1, y2, y3);
                    if (m > 1) {
                        const int tws = pp * s * tmul;
                        y1 = cmul(y1, s_tw[tws]); y2 = cmul(y2, s_tw[2 * tws]); y3 = cmul(y3, s_tw[3 * tws]);
                    }
                    dst[gph(bo)] = y0; dst[gph(bo + s)] = y1; dst[gph(bo + 2 * s)] = y2; dst[gph(bo + 3 * s)] = y3;
                }
            } else if (r == 2) {
                for (int bfly = lane; bfly < Nf / 2; bfly += 32) {
                    const int pp = sh >= 0 ? bfly >> sh : bfly / s, q = bfly - pp * s;
                    const float2 a0 = src[gph(q + s * pp)], a1 = src[gph(q + s * (pp + m))];
                    const int bo = q + s * 2 * pp;
                    dst[gph(bo)] = cadd(a0, a1);
                    dst[gph(bo + s)] = cmul(csub(a0, a1), s_tw[pp * s * tmul]);
                }
            } else if (r == 3) {
                for (int bfly = lane; bfly < Nf / 3; bfly += 32) {
                    const int pp = sh >= 0 ? bfly >> sh : bfly / s, q = bfly - pp * s;
                    const int bi = q + s * pp, bo = q + s * 3 * pp, tws = pp * s * tmul;
                    float2 a[3], y[3];
#pragma unroll
                    for (int i = 0; i < 3; ++i) a[i] = src[gph(bi + sm * i)];
                    gdft_small<3>(a, y);
                    dst[gph(bo)] = y[0];
#pragma unroll
                    for (int j = 1; j < 3; ++j) dst[gph(bo + s * j)] = cmul(y[j], s_tw[j * tws]);
                }
            } else if (r == 5) {
                for (int bfly = lane; bfly < Nf / 5; bfly += 32) {
                    const int pp = sh >= 0 ? bfly >> sh : bfly / s, q = bfly - pp * s;
                    const int bi = q + s * pp, bo = q + s * 5 * pp, tws = pp * s * tmul;
                    float2 a[5], y[5];
#pragma unroll
                    for (int i = 0; i < 5; ++i) a[i] = src[gph(bi + sm * i)];
                    gdft_small<5>(a, y);
                    dst[gph(bo)] = y[0];
#pragma unroll
                    for (int j = 1; j < 5; ++j) dst[gph(bo + s * j)] = cmul(y[j], s_tw[j * tws]);
                }
            } else {   // any other prime factor: one (butterfly, output) pair per work item
                const int wr = N / r;   // W_r^e = W_N^(e N/r)
                for (int e = lane; e < Nf; e += 32) {
                    const int bfly = e / r, j = e - bfly * r;
                    const int pp = sh >= 0 ? bfly >> sh : bfly / s, q = bfly - pp * s;
                    const int bi = q + s * pp;
                    float2 acc = src[gph(bi)];
                    int idx = 0;
                    for (int i = 1; i < r; ++i) {
                        idx += j;
                        if (idx >= r) idx -= r;
                        const float2 v = src[gph(bi + sm * i)], w = s_tw[idx * wr];
                        acc.x = fmaf(v.x, w.x, fmaf(-v.y, w.y, acc.x));
                        acc.y = fmaf(v.x, w.y, fmaf(v.y, w.x, acc.y));
                    }
                    dst[gph(q + s * (r * pp + j))] = cmul(acc, s_tw[pp * j * s * tmul]);
                }
            }
            __syncwarp();
            float2* t = src; src = dst; dst = t;
            ncur = m; s *= r;
        }

        // ------------------------------------------------------------------ power (or magnitude) of bins 0..N/2 -> pw[k]
        float* pw = reinterpret_cast<float*>(dst);
        if (packed) {
            // X[k] = E[k] + W_N^k O[k] and X[Nf - k] = conj(E[k] - W_N^k O[k]) share Z[k], Z[Nf - k] and the twiddle: one lane forms
            // both bins (k = 1 .. Nf/2; k = Nf/2 pairs with itself and is written twice); DC and Nyquist come from Z[0]
            for (int k = 1 + lane; k <= Nf / 2; k += 32) {
                const float2 zk = src[gph(k)], zm = src[gph(Nf - k)];
                const float er = 0.5f * (zk.x + zm.x), ei = 0.5f * (zk.y - zm.y);
                const float orr = 0.5f * (zk.y + zm.y), oi = -0.5f * (zk.x - zm.x);
                const float2 w = s_tw[k];
                const float tr = orr * w.x - oi * w.y, ti = fmaf(orr, w.y, oi * w.x);
                const float ar = er + tr, ai = ei + ti, br = er - tr, bi = ei - ti;
                float ea = fmaf(ar, ar, ai * ai), eb = fmaf(br, br, bi * bi);
                if (!g.use_power) { ea = sqrtf(ea); eb = sqrtf(eb); }   // src/fbank.rs:197-203
                pw[k] = ea;
                pw[Nf - k] = eb;
            }
            if (lane == 0) {
                const float2 z0 = src[0];
                const float x0 = z0.x + z0.y, xn = z0.x - z0.y;
                pw[0] = g.use_power ? x0 * x0 : fabsf(x0);
                pw[nb] = g.use_power ? xn * xn : fabsf(xn);
            }
        } else {
            for (int k = lane; k <= nb; k += 32) {
                const float2 z = src[gph(k)];
                float e = fmaf(z.x, z.x, z.y * z.y);
                if (!g.use_power) e = sqrtf(e);
                pw[k] = e;
            }
        }
        if (wsm) for (int i = lane; i < kpad; i += 32) pw[nb + 1 + i] = 0.f;   // rows a padded (zero-weight) entry may touch
        __syncwarp();

        // ------------------------------------------------------------------ banded projection + log + stores
        float v[kMaxMpl];
        float mx = -INFINITY;
#pragma unroll
        for (int sl = 0; sl < kMaxMpl; ++sl) {
            const int mrow = lane + 32 * sl;
            v[sl] = -INFINITY;
            if (mrow < p.n_mels) {
                float e = 0.f;
                if (wsm) {   // (the branch is warp-uniform; lanes past n_mels skip the slot)
                    int kb = 0;
                    for (int q = 0; q < sl; ++q) kb += g.kmax[q];
                    const float* wp = s_wts + 32 * kb + lane;
                    const float* pp = pw + __ldg(g.starts_t + mrow);
                    const int K = g.kmax[sl];
#pragma unroll 4
                    for (int i = 0; i < K; ++i) e = fmaf(wp[32 * i], pp[i], e);
                } else {
                    const int b0 = __ldg(g.bands + 3 * mrow), cnt = __ldg(g.bands + 3 * mrow + 1), wo = __ldg(g.bands + 3 * mrow + 2);
                    for (int i = 0; i < cnt; ++i) e = fmaf(__ldg(g.weights + wo + i), pw[b0 + i], e);
                }
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
