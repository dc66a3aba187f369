// melspec_generic2.cuh — the general plan, two frames per warp ("pair form").
//
// Same algorithm and tables as melspec_generic.cuh (mixed-radix Stockham FFT of the frame's even / odd samples in the warp's
// private shared-memory ping-pong buffers, power, CSR banded projection, log, stores), restated so that a warp carries TWO
// neighbouring frames of a clip (2j, 2j+1) through it at once as the two halves of Blackwell's packed FADD2 / FMUL2 / FFMA2
// instructions: a buffer element is a float4 (re A, re B, im A, im B), every butterfly, twiddle product, untangle step and
// projection FMA advances both frames with one instruction, and all index arithmetic, twiddle loads and loop control are paid
// once per pair.  Unlike the two-real-frames-per-complex-transform packing of the specialised kernels the two frames never mix:
// each is its own transform, so every frame's rounding noise stays relative to its own level.
//
// NFT > 0 compiles the transform length (Nf = NFT complex points, fft_size = 2 NFT) and with it the radix schedule, all strides
// and all trip counts into the kernel (loops fully unrolled, index arithmetic folded); NFT = 0 reads them from GParams at run
// time (any even or odd size).  The host instantiates the sizes the reference's users meet outside the two specialised plans
// (fft 256, 480, 512 with other hops, 640, 800, 1024, 2048) and falls back to NFT = 0 — or, where two frames' buffers leave
// too few warps per SM, to the one-frame kernel of melspec_generic.cuh.
#pragma once
#include "melspec_generic.cuh"

namespace melspec {

// element i of a pair buffer lives at i + (i >> 3): 128-bit accesses are served per quarter-warp (8 lanes x 16 bytes), and with
// one pad slot per 8 elements the stride-8 stores of the first radix-8 stage (9 = 1 mod 8) and the unit-stride accesses of the
// later ones both touch 8 different 16-byte bank groups
__device__ __forceinline__ int gph4(int i) { return i + (i >> 3); }
__host__ __device__ inline int generic2_buf_elems(int nf) { return nf + (nf >> 3) + 2; }

struct cpair { f2 re, im; };   // one complex value of frame A (.x) and of frame B (.y)
__device__ __forceinline__ cpair ldp(const float4* b, int i) { const float4 v = b[gph4(i)]; return cpair{make_float2(v.x, v.y), make_float2(v.z, v.w)}; }
__device__ __forceinline__ void stp(float4* b, int i, const cpair v) { b[gph4(i)] = make_float4(v.re.x, v.re.y, v.im.x, v.im.y); }
__device__ __forceinline__ cpair padd(const cpair a, const cpair b) { return cpair{add2(a.re, b.re), add2(a.im, b.im)}; }
__device__ __forceinline__ cpair psub(const cpair a, const cpair b) { return cpair{sub2(a.re, b.re), sub2(a.im, b.im)}; }
// a * w, the same twiddle for both frames: 2 FMUL2 + 2 FFMA2
__device__ __forceinline__ cpair pmul(const cpair a, const float2 w) {
    return cpair{fma2c(-w.y, a.im, mul2c(w.x, a.re)), fma2c(w.y, a.re, mul2c(w.x, a.im))};
}

__device__ __forceinline__ void pdft4(const cpair b0, const cpair b1, const cpair b2, const cpair b3, cpair& y0, cpair& y1, cpair& y2, cpair& y3) {
    const cpair s02 = padd(b0, b2), d02 = psub(b0, b2), s13 = padd(b1, b3), d13 = psub(b1, b3);
    y0 = padd(s02, s13);
    y2 = psub(s02, s13);
    y1 = cpair{add2(d02.re, d13.im), sub2(d02.im, d13.re)};   // d02 - i d13
    y3 = cpair{sub2(d02.re, d13.im), add2(d02.im, d13.re)};   // d02 + i d13
}
__device__ __forceinline__ void pdft8(cpair (&a)[8]) {
    constexpr float c = 0.70710678118654752f;
    const cpair t0 = padd(a[0], a[4]), t4 = psub(a[0], a[4]);
    const cpair t1 = padd(a[1], a[5]), d1 = psub(a[1], a[5]);
    const cpair t2 = padd(a[2], a[6]), d2 = psub(a[2], a[6]);
    const cpair t3 = padd(a[3], a[7]), d3 = psub(a[3], a[7]);
    const cpair t5 = cpair{mul2c(c, add2(d1.re, d1.im)), mul2c(c, sub2(d1.im, d1.re))};     // d1 W_8
    const cpair t6 = cpair{d2.im, make_float2(-d2.re.x, -d2.re.y)};                          // -i d2
    const cpair t7 = cpair{mul2c(c, sub2(d3.im, d3.re)), mul2c(-c, add2(d3.re, d3.im))};    // d3 W_8^3
    pdft4(t0, t1, t2, t3, a[0], a[2], a[4], a[6]);
    pdft4(t4, t5, t6, t7, a[1], a[3], a[5], a[7]);
}
template <int R>
__device__ __forceinline__ void pdft_small(const cpair (&a)[R], cpair (&y)[R]) {
#pragma unroll
    for (int j = 0; j < R; ++j) {
        cpair acc = a[0];
#pragma unroll
        for (int i = 1; i < R; ++i) {
            const int e = (i * j) % R;                       // W_R^e = (cos, -sin)(2 pi e / R)
            const float wc = groot_cos<R>(e), ws = -groot_sin<R>(e);
            acc.re = fma2c(wc, a[i].re, fma2c(-ws, a[i].im, acc.re));
            acc.im = fma2c(ws, a[i].re, fma2c(wc, a[i].im, acc.im));
        }
        y[j] = acc;
    }
}

// One Stockham stage of radix R on sub-transforms of length n = R m, stride s (see melspec_generic.cuh):
//   y[q + s (R p + j)] = W_Nf^(p j s) * sum_i x[q + s (p + m i)] W_R^(i j)
// Every argument is a compile-time constant in the NFT > 0 instantiations (after inlining), a run-time value otherwise.
// stw (compile-time sizes): this stage's own twiddle table, W^(j p s) at stw[(j - 1) m + p] — consecutive lanes read consecutive (or
// the same) entries, where the strided reads of the shared W_N^k table collide up to eightfold, and nothing is derived by arithmetic.
template <int R>
__device__ __forceinline__ void pstage(const float4* src, float4* dst, const float2* s_tw, const float2* stw, const int Nf, const int m,
                                       const int s, const int sh, const int tmul, const int lane) {
    const int cnt = Nf / R, sm = s * m, iters = (cnt + 31) >> 5;
#pragma unroll
    for (int it = 0; it < iters; ++it) {
        const int bfly = lane + 32 * it;
        if (bfly < cnt) {
            const int pp = sh >= 0 ? bfly >> sh : bfly / s, q = bfly - pp * s;
            const int bi = q + s * pp, bo = q + s * R * pp, tws = pp * s * tmul;
            if (R == 8) {
                cpair a[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = ldp(src, bi + sm * i);
                pdft8(a);
                if (m > 1 && stw != nullptr) {
#pragma unroll
                    for (int j = 1; j < 8; ++j) a[j] = pmul(a[j], stw[(j - 1) * m + pp]);
                } else if (m > 1) {   // W^(j p s), j = 1..7, from three table reads
                    const float2 w1 = s_tw[tws], w2 = s_tw[2 * tws], w4 = s_tw[4 * tws];
                    const float2 w3 = cmul(w1, w2), w5 = cmul(w4, w1), w6 = cmul(w4, w2), w7 = cmul(w4, w3);
                    a[1] = pmul(a[1], w1); a[2] = pmul(a[2], w2); a[3] = pmul(a[3], w3); a[4] = pmul(a[4], w4);
                    a[5] = pmul(a[5], w5); a[6] = pmul(a[6], w6); a[7] = pmul(a[7], w7);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) stp(dst, bo + s * j, a[j]);
            } else if (R == 4) {
                cpair y0, y1, y2, y3;
                pdft4(ldp(src, bi), ldp(src, bi + sm), ldp(src, bi + 2 * sm), ldp(src, bi + 3 * sm), y0, y1, y2, y3);
                if (m > 1 && stw != nullptr) { y1 = pmul(y1, stw[pp]); y2 = pmul(y2, stw[m + pp]); y3 = pmul(y3, stw[2 * m + pp]); }
                else if (m > 1) { y1 = pmul(y1, s_tw[tws]); y2 = pmul(y2, s_tw[2 * tws]); y3 = pmul(y3, s_tw[3 * tws]); }
                stp(dst, bo, y0); stp(dst, bo + s, y1); stp(dst, bo + 2 * s, y2); stp(dst, bo + 3 * s, y3);
            } else if (R == 2) {
                const cpair a0 = ldp(src, bi), a1 = ldp(src, bi + sm);
                stp(dst, bo, padd(a0, a1));
                const cpair d = psub(a0, a1);
                stp(dst, bo + s, m > 1 ? pmul(d, stw != nullptr ? stw[pp] : s_tw[tws]) : d);
            } else {   // R = 3, 5
                cpair a[R], y[R];
#pragma unroll
                for (int i = 0; i < R; ++i) a[i] = ldp(src, bi + sm * i);
                pdft_small<R>(a, y);
                stp(dst, bo, y[0]);
#pragma unroll
                for (int j = 1; j < R; ++j) stp(dst, bo + s * j, m > 1 ? pmul(y[j], stw != nullptr ? stw[(j - 1) * m + pp] : s_tw[j * tws]) : y[j]);
            }
        }
    }
}

// any other prime factor r: one (butterfly, output) pair per work item, an r-term sum per output
__device__ __forceinline__ void pstage_prime(const float4* src, float4* dst, const float2* s_tw, const int N, const int Nf, const int r,
                                             const int m, const int s, const int sh, const int tmul, const int lane) {
    const int wr = N / r, sm = s * m;   // W_r^e = W_N^(e N/r)
    for (int e = lane; e < Nf; e += 32) {
        const int bfly = e / r, j = e - bfly * r;
        const int pp = sh >= 0 ? bfly >> sh : bfly / s, q = bfly - pp * s;
        const int bi = q + s * pp;
        cpair acc = ldp(src, bi);
        int idx = 0;
        for (int i = 1; i < r; ++i) {
            idx += j;
            if (idx >= r) idx -= r;
            const cpair v = ldp(src, bi + sm * i);
            const float2 w = s_tw[idx * wr];
            acc.re = fma2c(w.x, v.re, fma2c(-w.y, v.im, acc.re));
            acc.im = fma2c(w.y, v.re, fma2c(w.x, v.im, acc.im));
        }
        stp(dst, q + s * (r * pp + j), pmul(acc, s_tw[pp * j * s * tmul]));
    }
}

__host__ __device__ constexpr int g2_log2_or_neg(int s) {
    int sh = 0;
    while ((1 << sh) < s) ++sh;
    return (1 << sh) == s ? sh : -1;
}
// compile-time schedule: 8s, 4s, a 2, then 3s and 5s (the order build_tables_generic uses)
__host__ __device__ constexpr int g2_radix(int ncur) { return ncur % 8 == 0 ? 8 : ncur % 4 == 0 ? 4 : ncur % 2 == 0 ? 2 : ncur % 3 == 0 ? 3 : 5; }
// entries of the per-stage twiddle tables of a compile-time size: (R - 1) m for every stage with m > 1
__host__ __device__ constexpr int g2_stw_elems(int nf) {
    int tot = 0, ncur = nf;
    while (ncur > 1) {
        const int r = g2_radix(ncur), m = ncur / r;
        if (m > 1) tot += (r - 1) * m;
        ncur = m;
    }
    return tot;
}
template <int NF, int NCUR, int S, int OFF>
__device__ __forceinline__ void pstages_ct(float4*& src, float4*& dst, const float2* s_tw, const float2* s_stw, const int lane) {
    if constexpr (NCUR > 1) {
        constexpr int R = g2_radix(NCUR), M = NCUR / R;
        static_assert(NCUR % R == 0, "compile-time sizes are products of 2, 3 and 5");
        pstage<R>(src, dst, s_tw, s_stw + OFF, NF, M, S, g2_log2_or_neg(S), 2, lane);
        __syncwarp();
        float4* t = src; src = dst; dst = t;
        pstages_ct<NF, M, S * R, OFF + (M > 1 ? (R - 1) * M : 0)>(src, dst, s_tw, s_stw, lane);
    }
}
// In-place form of a compile-time stage (one buffer per warp instead of two: more warps fit on the SM).  A Stockham stage writes where
// other lanes read, so every lane first pulls all of its butterflies into registers (NF / 8 registers) and transforms them; after a
// warp barrier the results go back into the same buffer.
template <int R, int NF, int M, int S>
__device__ __forceinline__ void pstage_inplace(float4* buf, const float2* stw, const int lane) {
    constexpr int CNT = NF / R, SM = S * M, ITERS = (CNT + 31) / 32, SH = g2_log2_or_neg(S);
    cpair a[ITERS][R];
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int bfly = lane + 32 * it;
        if (bfly < CNT) {
            const int pp = SH >= 0 ? bfly >> SH : bfly / S, q = bfly - pp * S, bi = q + S * pp;
#pragma unroll
            for (int i = 0; i < R; ++i) a[it][i] = ldp(buf, bi + SM * i);
            if constexpr (R == 8) pdft8(a[it]);
            else if constexpr (R == 4) { cpair y0, y1, y2, y3; pdft4(a[it][0], a[it][1], a[it][2], a[it][3], y0, y1, y2, y3); a[it][0] = y0; a[it][1] = y1; a[it][2] = y2; a[it][3] = y3; }
            else if constexpr (R == 2) { const cpair t = padd(a[it][0], a[it][1]); a[it][1] = psub(a[it][0], a[it][1]); a[it][0] = t; }
            else {
                cpair y[R];
                pdft_small<R>(a[it], y);
#pragma unroll
                for (int j = 0; j < R; ++j) a[it][j] = y[j];
            }
            if (M > 1) {
#pragma unroll
                for (int j = 1; j < R; ++j) a[it][j] = pmul(a[it][j], stw[(j - 1) * M + pp]);
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int bfly = lane + 32 * it;
        if (bfly < CNT) {
            const int pp = SH >= 0 ? bfly >> SH : bfly / S, q = bfly - pp * S, bo = q + S * R * pp;
#pragma unroll
            for (int j = 0; j < R; ++j) stp(buf, bo + S * j, a[it][j]);
        }
    }
}
template <int NF, int NCUR, int S, int OFF>
__device__ __forceinline__ void pstages_inplace_ct(float4* buf, const float2* s_stw, const int lane) {
    if constexpr (NCUR > 1) {
        constexpr int R = g2_radix(NCUR), M = NCUR / R;
        static_assert(NCUR % R == 0, "compile-time sizes are products of 2, 3 and 5");
        pstage_inplace<R, NF, M, S>(buf, s_stw + OFF, lane);
        __syncwarp();
        pstages_inplace_ct<NF, M, S * R, OFF + (M > 1 ? (R - 1) * M : 0)>(buf, s_stw, lane);
    }
}
// fill the stage tables from the W_N^k table in global memory (once per CTA): W_Nf^(j p s) = W_N^(2 j p s), and j p s < Nf
template <int NF, int NCUR, int S, int OFF>
__device__ __forceinline__ void fill_stw_ct(const float2* tw, float2* s_stw) {
    if constexpr (NCUR > 1) {
        constexpr int R = g2_radix(NCUR), M = NCUR / R;
        if constexpr (M > 1) {
            for (int i = threadIdx.x; i < (R - 1) * M; i += blockDim.x) {
                const int j = i / M + 1, pp = i - (j - 1) * M;
                s_stw[OFF + i] = __ldg(tw + 2 * j * pp * S);
            }
        }
        fill_stw_ct<NF, M, S * R, OFF + (M > 1 ? (R - 1) * M : 0)>(tw, s_stw);
    }
}

// LB256x3: compiled for at most 256 threads per CTA and three CTAs per SM (80 registers per thread instead of up to 128): small
// transforms are latency-bound and register-limited, so more resident warps can beat a few spills.
template <int NFT, bool INPLACE = false, bool LB256x3 = false>
__global__ void __launch_bounds__(LB256x3 ? 256 : 512, LB256x3 ? 3 : 1) melspec_generic_pair_kernel(const KParams p, const GParams g) {
    static_assert(!INPLACE || NFT != 0, "the one-buffer form needs a compile-time size");
    extern __shared__ __align__(16) unsigned char gsm[];
    const int N = NFT ? 2 * NFT : g.N, Nf = NFT ? NFT : g.Nf, tmul = N / Nf;   // W_Nf^k = W_N^(tmul k)
    const bool packed = NFT ? true : Nf != N;
    const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2* s_tw = reinterpret_cast<float2*>(gsm);
    const int nfp = generic2_buf_elems(Nf);
    // [twiddles | transposed band weights | per-warp ping-pong buffers], each part 16-byte aligned.  The projection reads its weights
    // from shared memory: with most of the L1 carved out for the buffers, global reads of the table miss L1 behind the streaming PCM
    // and every batch of the loop waits a full L2 round trip (measured: 30 % of all stall samples)
    // (compile-time sizes keep only W_N^k, k <= N/4, for the untangle step: their stages have their own tables)
    const int tw_elems = NFT ? N / 4 + 1 : N;
    const size_t tw_bytes = ((size_t)8 * tw_elems + 15) & ~(size_t)15, w_bytes = ((size_t)4 * g.n_weights_t + 15) & ~(size_t)15;
    constexpr size_t stw_bytes = ((size_t)8 * g2_stw_elems(NFT ? NFT : 1) + 15) & ~(size_t)15;   // per-stage twiddle tables (compile-time sizes)
    float* s_wts = reinterpret_cast<float*>(gsm + tw_bytes);
    float2* s_stw = reinterpret_cast<float2*>(gsm + tw_bytes + w_bytes);
    float4* buf0 = reinterpret_cast<float4*>(gsm + tw_bytes + w_bytes + stw_bytes) + (size_t)warp * (INPLACE ? 1 : 2) * nfp;
    float4* buf1 = INPLACE ? buf0 : buf0 + nfp;
    for (int i = threadIdx.x; i < tw_elems; i += blockDim.x) s_tw[i] = g.tw[i];
    for (int i = threadIdx.x; i < g.n_weights_t; i += blockDim.x) s_wts[i] = g.weights_t[i];
    if constexpr (NFT != 0) fill_stw_ct<NFT, NFT, 1, 0>(g.tw, s_stw);
    __syncthreads();

    const int L = p.frame_len;
    const float inv_len = 1.0f / (float)L;
    const int nb = N / 2;   // last bin formed
    const int ppc = (p.frames_per_clip + 1) >> 1;   // pairs per clip
    const long long n_pairs = (long long)ppc * p.n_clips;

    // this lane's mel rows: first bin of the band (the weights of slot sl are s_wts[32 (kbase[sl] + i) + lane], zero past the band)
    int bnd0[kMaxMpl], kbase[kMaxMpl], kpad = 0;
#pragma unroll
    for (int sl = 0; sl < kMaxMpl; ++sl) {
        const int mrow = lane + 32 * sl;
        bnd0[sl] = mrow < p.n_mels ? __ldg(g.starts_t + mrow) : 0;
        kbase[sl] = sl == 0 ? 0 : kbase[sl - 1] + g.kmax[sl - 1];
        kpad = max(kpad, g.kmax[sl]);
    }

    for (long long u = (long long)blockIdx.x * nw + warp; u < n_pairs; u += (long long)gridDim.x * nw) {
        const int clip = (int)(u / ppc);
        const int f0 = 2 * (int)(u - (long long)clip * ppc);
        int len = p.n_samples;
        int nfr = p.frames_per_clip;   // frames of this clip that are computed
        bool zero_fill = false;         // NeMo ragged batch: columns past the clip's own frame count are written as zeros
        if (p.lens) {
            len = min(p.lens[clip], p.n_samples);
            if (g.mode != 2) nfr = min(nfr, len < L ? 0 : (len - L) / p.hop + 1);
            else { nfr = min(nfr, len <= 0 ? 0 : g.center ? len / p.hop + 1 : (len < N ? 0 : (len - N) / p.hop + 1)); zero_fill = true; }
        }
        const bool va = f0 < nfr, vb = f0 + 1 < nfr;
        float* oc = p.out + (long long)clip * p.out_clip_stride;
        if (zero_fill && !vb) {   // (src/mel.rs:336, 387-395)
            for (int q = va ? 1 : 0; q < 2; ++q) {
                const int f = f0 + q;
                if (f < p.frames_per_clip)
                    for (int mrow = lane; mrow < p.n_mels; mrow += 32) {
                        if (p.layout == 0) oc[(long long)f * p.n_mels + mrow] = 0.f;
                        else oc[(long long)mrow * p.out_row_stride + f] = 0.f;
                    }
            }
        }
        if (!va) continue;   // frames past a short clip's own frame count are left untouched (warp-uniform)
        const float* x = p.pcm + (long long)clip * p.clip_stride;
        const long long sa = (long long)f0 * p.hop + p.frame_offset;
        const long long sb = sa + p.hop;
        auto at = [&](long long i) -> float { return (i >= 0 && i < len) ? __ldg(x + i) : 0.f; };

        // ------------------------------------------------------------------ prologue: windowed sample n of both frames
        // interior pair (all but the first and last few of a clip): every sample the prologue touches, including the look-back
        // sample of the pre-emphasis, exists, so the loads need no bounds checks
        const bool interior = vb && sa >= 1 && sb + (long long)N <= len;
        const float* pa = x + sa;
        const float* pb = x + sb;
        float mua = 0.f, mub = 0.f;
        if (g.mode == 1) {   // per-frame mean (src/fbank.rs:166-170)
            float s0 = 0.f, s1 = 0.f;
            if (interior) for (int n = lane; n < L; n += 32) { s0 += __ldg(pa + n); s1 += __ldg(pb + n); }
            else for (int n = lane; n < L; n += 32) { s0 += at(sa + n); s1 += at(sb + n); }
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
            mua = s0 * inv_len; mub = s1 * inv_len;
        }
        auto sample = [&](const long long s0, const float mu, const int n) -> float {
            if (n >= L) return 0.f;   // zero padding up to the FFT size
            const float w = __ldg(g.window + n);
            if (g.mode == 0) return at(s0 + n) * w;
            if (g.mode == 1) {        // src/fbank.rs:172-190: frame 0 of a clip has no look-back sample
                float a = at(s0 + n) - mu;
                if (n > 0 || s0 > 0) a = fmaf(-p.preemph, at(s0 + n - 1) - mu, a);
                return a * w;
            }
            const long long ia = s0 + n;   // src/mel.rs:696-706: wave[i] = x[i] - c x[i-1] (i >= 1), zero outside the clip
            return (ia >= 0 && ia < len) ? fmaf(-p.preemph, at(ia - 1), at(ia)) * w : 0.f;
        };
        if (packed && g.mode == 0 && g.vec2 && sb + N <= len) {   // both frames inside the clip, 8-byte aligned pairs
            const float2* xa = reinterpret_cast<const float2*>(x + sa);
            const float2* xb = reinterpret_cast<const float2*>(x + sb);
            const float2* w2 = reinterpret_cast<const float2*>(g.window);
            const int iters = (Nf + 31) >> 5;   // (a compile-time trip count when NFT > 0)
#pragma unroll
            for (int it = 0; it < iters; ++it) {
                const int n = lane + 32 * it;
                if (n < Nf) {
                    const float2 a = __ldg(xa + n), b = __ldg(xb + n), w = __ldg(w2 + n);
                    buf0[gph4(n)] = make_float4(a.x * w.x, b.x * w.x, a.y * w.y, b.y * w.y);
                }
            }
        } else if (interior && packed && g.vec2 && (L & 1) == 0) {
            // Kaldi / NeMo, interior pair, 8-byte aligned: one 64-bit load per frame and element; the look-back sample of the
            // pre-emphasis is the neighbouring lane's second sample (lane 0 loads its own)
            const float2* xa = reinterpret_cast<const float2*>(pa);
            const float2* xb = reinterpret_cast<const float2*>(pb);
            const float2* w2 = reinterpret_cast<const float2*>(g.window);
            const int iters = (Nf + 31) >> 5;
#pragma unroll(NFT ? 4 : 1)
            for (int it = 0; it < iters; ++it) {
                const int n = lane + 32 * it;
                const bool live = n < Nf && 2 * n < L;
                float2 a = make_float2(0.f, 0.f), b = a, w = a;
                if (live) { a = __ldg(xa + n); b = __ldg(xb + n); w = __ldg(w2 + n); }
                float am1 = __shfl_up_sync(0xffffffffu, a.y, 1), bm1 = __shfl_up_sync(0xffffffffu, b.y, 1);
                if (lane == 0 && live) { am1 = __ldg(pa + 2 * n - 1); bm1 = __ldg(pb + 2 * n - 1); }
                if (g.mode == 1) { a.x -= mua; a.y -= mua; am1 -= mua; b.x -= mub; b.y -= mub; bm1 -= mub; }   // src/fbank.rs:172-190
                if (n < Nf)   // y[i] = x[i] - c x[i-1], windowed (src/fbank.rs:172-190, src/mel.rs:696-706); zero padding past L
                    buf0[gph4(n)] = make_float4(fmaf(-p.preemph, am1, a.x) * w.x, fmaf(-p.preemph, bm1, b.x) * w.x,
                                                fmaf(-p.preemph, a.x, a.y) * w.y, fmaf(-p.preemph, b.x, b.y) * w.y);
            }
        } else if (interior) {
            auto fast = [&](const float* q, const float mu, const int n) -> float {
                if (n >= L) return 0.f;   // zero padding up to the FFT size
                const float w = __ldg(g.window + n);
                if (g.mode == 0) return __ldg(q + n) * w;
                if (g.mode == 1) return fmaf(-p.preemph, __ldg(q + n - 1) - mu, __ldg(q + n) - mu) * w;   // src/fbank.rs:172-190
                return fmaf(-p.preemph, __ldg(q + n - 1), __ldg(q + n)) * w;                             // src/mel.rs:696-706
            };
            const int iters = (Nf + 31) >> 5;
#pragma unroll(NFT ? 4 : 1)
            for (int it = 0; it < iters; ++it) {
                const int n = lane + 32 * it;
                if (n < Nf)
                    buf0[gph4(n)] = packed ? make_float4(fast(pa, mua, 2 * n), fast(pb, mub, 2 * n), fast(pa, mua, 2 * n + 1), fast(pb, mub, 2 * n + 1))
                                           : make_float4(fast(pa, mua, n), fast(pb, mub, n), 0.f, 0.f);
            }
        } else if (packed) {
            for (int n = lane; n < Nf; n += 32)
                buf0[gph4(n)] = make_float4(sample(sa, mua, 2 * n), vb ? sample(sb, mub, 2 * n) : 0.f, sample(sa, mua, 2 * n + 1),
                                            vb ? sample(sb, mub, 2 * n + 1) : 0.f);
        } else {
            for (int n = lane; n < Nf; n += 32) buf0[gph4(n)] = make_float4(sample(sa, mua, n), vb ? sample(sb, mub, n) : 0.f, 0.f, 0.f);
        }
        __syncwarp();

        // ------------------------------------------------------------------ Stockham autosort FFT of both frames
        float4* src = buf0;
        float4* dst = buf1;
        if constexpr (INPLACE) {
            pstages_inplace_ct<NFT, NFT, 1, 0>(buf0, s_stw, lane);
        } else if constexpr (NFT != 0) {
            pstages_ct<NFT, NFT, 1, 0>(src, dst, s_tw, s_stw, lane);
        } else {
            int ncur = Nf, s = 1;
            for (int st = 0; st < g.n_stages; ++st) {
                const int r = g.radix[st], m = ncur / r, sh = g.sshift[st];
                if (r == 8) pstage<8>(src, dst, s_tw, nullptr, Nf, m, s, sh, tmul, lane);
                else if (r == 4) pstage<4>(src, dst, s_tw, nullptr, Nf, m, s, sh, tmul, lane);
                else if (r == 2) pstage<2>(src, dst, s_tw, nullptr, Nf, m, s, sh, tmul, lane);
                else if (r == 3) pstage<3>(src, dst, s_tw, nullptr, Nf, m, s, sh, tmul, lane);
                else if (r == 5) pstage<5>(src, dst, s_tw, nullptr, Nf, m, s, sh, tmul, lane);
                else pstage_prime(src, dst, s_tw, N, Nf, r, m, s, sh, tmul, lane);
                __syncwarp();
                float4* t = src; src = dst; dst = t;
                ncur = m; s *= r;
            }
        }

        // ------------------------------------------------------------------ power (or magnitude) of bins 0..N/2 -> pw[k] = (A, B)
        float2* pw = reinterpret_cast<float2*>(dst);
        auto mag = [&](const f2 xr, const f2 xi) -> f2 {
            f2 e = fma2(xr, xr, mul2(xi, xi));
            if (!g.use_power) e = make_float2(sqrtf(e.x), sqrtf(e.y));   // src/fbank.rs:197-203
            return e;
        };
        // X[k] = E[k] + W_N^k O[k] and X[Nf - k] = conj(E[k] - W_N^k O[k]) come from the same two values Z[k], Z[Nf - k] and
        // the same twiddle: one lane forms both bins (k = 1 .. Nf/2; k = Nf/2 pairs with itself and is written twice)
        auto bins2 = [&](const int k, const cpair zk, const cpair zm) {
            const f2 er = mul2c(0.5f, add2(zk.re, zm.re)), ei = mul2c(0.5f, sub2(zk.im, zm.im));
            const f2 orr = mul2c(0.5f, add2(zk.im, zm.im)), oi = mul2c(-0.5f, sub2(zk.re, zm.re));
            const float2 w = s_tw[k];
            const f2 tr = fma2c(-w.y, oi, mul2c(w.x, orr)), ti = fma2c(w.y, orr, mul2c(w.x, oi));
            pw[k] = mag(add2(er, tr), add2(ei, ti));
            pw[Nf - k] = mag(sub2(er, tr), sub2(ei, ti));
        };
        auto bins_dc = [&](const cpair z0) {   // DC and Nyquist: X[0] = Re Z[0] + Im Z[0], X[N/2] = Re Z[0] - Im Z[0]
            const f2 x0 = add2(z0.re, z0.im), xn = sub2(z0.re, z0.im);
            pw[0] = g.use_power ? mul2(x0, x0) : make_float2(fabsf(x0.x), fabsf(x0.y));
            pw[nb] = g.use_power ? mul2(xn, xn) : make_float2(fabsf(xn.x), fabsf(xn.y));
        };
        if constexpr (INPLACE) {   // the powers overwrite the spectrum: all of it is read into registers first
            constexpr int CNT = NFT / 2, ITERS = (CNT + 31) / 32;
            cpair zk[ITERS], zm[ITERS];
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                const int k = 1 + lane + 32 * it;
                if (k <= CNT) { zk[it] = ldp(src, k); zm[it] = ldp(src, NFT - k); }
            }
            const cpair z0 = ldp(src, 0);
            __syncwarp();
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                const int k = 1 + lane + 32 * it;
                if (k <= CNT) bins2(k, zk[it], zm[it]);
            }
            if (lane == 0) bins_dc(z0);
        } else if (packed) {
            const int cnt = Nf / 2, iters = (cnt + 31) >> 5;
#pragma unroll
            for (int it = 0; it < iters; ++it) {
                const int k = 1 + lane + 32 * it;
                if (k <= cnt) bins2(k, ldp(src, k), ldp(src, Nf - k));
            }
            if (lane == 0) bins_dc(ldp(src, 0));
        } else {
            for (int k = lane; k <= nb; k += 32) {
                const cpair z = ldp(src, k);
                pw[k] = mag(z.re, z.im);
            }
        }
        for (int i = lane; i < kpad; i += 32) pw[nb + 1 + i] = make_float2(0.f, 0.f);   // rows a padded (zero-weight) entry may touch
        __syncwarp();

        // ------------------------------------------------------------------ banded projection + log + stores
        f2 v[kMaxMpl];
        f2 mx = make_float2(-3.0e38f, -3.0e38f);
#pragma unroll
        for (int sl = 0; sl < kMaxMpl; ++sl) {
            const int mrow = lane + 32 * sl;
            v[sl] = make_float2(-3.0e38f, -3.0e38f);
            if (sl * 32 < p.n_mels) {   // (warp-uniform; lanes past n_mels carry zero weights and are not stored)
                const float* wp = s_wts + 32 * kbase[sl] + lane;
                const float2* pp2 = pw + bnd0[sl];
                f2 e = make_float2(0.f, 0.f);
                const int K = g.kmax[sl];
#pragma unroll 4
                for (int i = 0; i < K; ++i) e = fma2c(wp[32 * i], pp2[i], e);
                if (g.mode == 2) e = make_float2(logf(e.x + p.log_add), logf(e.y + p.log_add));          // ln(E + guard), src/mel.rs:365-368
                else if (g.mode == 1) {                                                               // max(E, floor), optional ln, src/fbank.rs:207-221
                    e = make_float2(fmaxf(e.x, p.floor_val), fmaxf(e.y, p.floor_val));
                    if (g.use_log) e = make_float2(logf(e.x), logf(e.y));
                } else e = make_float2(log10f(fmaxf(e.x, p.floor_val)), log10f(fmaxf(e.y, p.floor_val)));   // src/mel.rs:148-168
                v[sl] = e;
                if (mrow < p.n_mels) mx = make_float2(fmaxf(mx.x, e.x), fmaxf(mx.y, e.y));
            }
        }
        if (p.normalize) {   // per-frame clamp to max - 8, then (x + 4)/4 (src/mel.rs:645-654)
            mx = make_float2(warp_max_f32(mx.x) - 8.0f, warp_max_f32(mx.y) - 8.0f);
        }
#pragma unroll
        for (int sl = 0; sl < kMaxMpl; ++sl) {
            const int mrow = lane + 32 * sl;
            if (mrow < p.n_mels) {
                f2 y = v[sl];
                if (p.normalize) y = make_float2((fmaxf(y.x, mx.x) + 4.0f) * 0.25f, (fmaxf(y.y, mx.y) + 4.0f) * 0.25f);
                if (p.layout == 0) {
                    oc[(long long)f0 * p.n_mels + mrow] = y.x;
                    if (vb) oc[(long long)(f0 + 1) * p.n_mels + mrow] = y.y;
                } else {
                    float* r = oc + (long long)mrow * p.out_row_stride + f0;
                    r[0] = y.x;
                    if (vb) r[1] = y.y;
                }
            }
        }
        __syncwarp();   // the next pair overwrites both buffers
    }
}

}  // namespace melspec
