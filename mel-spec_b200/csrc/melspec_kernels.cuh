// melspec_b200 device code: the fused window -> FFT -> |X|^2 -> banded mel projection -> log -> normalise kernel.
//
// Written from scratch for sm_100a.  What it replaces in the reference (wavey-ai/mel-spec):
//   host windowing/framing   src/stft.rs:147-169      (here: Hann folded into the first FFT stage, in registers)
//   cufftExecZ2Z             src/cuda.rs:356-357      (here: 2 real frames per complex 400-point FFT, 20x20
//                                                      Cooley-Tukey with Good-Thomas 4x5 codelets in registers)
//   mel_kernel               src/cuda_kernels.cu:5-47 (here: sparse banded projection from shared memory)
//   norm_mel_vec on the host src/mel.rs:458-469       (here: fused, per frame)
// so PCM is read once from HBM (TMA bulk copies into shared memory) and mel frames are written once.
//
// Thread organisation ("plan 400"): a warp owns 3 complex FFTs = 6 frames per pass; 10 lanes cooperate on one FFT
// (lane = 10*g + t: t = worker 0..9, g = FFT 0..2; lanes 30,31 shadow lane 29).  N = 400 = 20 x 20:
//   step 1  worker t transforms columns n2 = 2t, 2t+1 (elements x[20*n1 + n2]) with a 20-point DFT -> Y[n2][k1]
//   exchange through the warp's private shared-memory slab Z[slot(k1)][g][n2 pair]   (only __syncwarp, no CTA barrier)
//   step 3  worker t owns rows k1 = t and 20-t (t = 0: rows 0 and 10): twiddle, 20-point DFT over n2 -> X[k1 + 20*k2].
//           The twiddles W_400^(t*n2) come from a shared table (registers in the 8-warp build); row 20-t uses their
//           conjugates and a rotation of the DFT
//           outputs by one (W_400^((20-t)n2) = W_20^n2 * conj W_400^(t*n2)); row 10 is pre-rotated by W_40^(-n2) when
//           it is written, which makes worker 0 (twiddle 1) follow exactly the same code.
//   untangle: frame A = Re, frame B = Im of the packed input, |A[k]|^2 = |Z[k] + conj Z[N-k]|^2 / 4 (the 1/4 lives
//   in the mel weights); both Z[k] and Z[N-k] sit in the same worker by construction (rows k1 and 20-k1).
//   projection: powers are stored in natural bin order (one plane per FFT), a mel band is a run of consecutive rows, so a
//   lane's entries are K consecutive rows from a host-chosen, conflict-free window start; weights-only table.
// Bins 1..200 are produced (DC never is: every supported filterbank has a zero DC column; the host checks).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace melspec {

// ------------------------------------------------------------------------------------------------ parameters
struct KParams {
    const float* pcm;        // [n_clips][clip_stride]
    float* out;              // frame-major [n_clips][F][n_mels] or mel-major [n_clips][n_mels][F]
    const int32_t* lens;     // optional per-clip valid samples
    long long clip_stride;   // samples
    long long out_clip_stride;  // floats
    int n_samples;           // samples per clip (copy limit)
    int frames_per_clip;     // F = num_frames(n_samples)
    int wtiles_per_clip;     // warp tiles (6 frames) per clip
    int n_wtiles;            // n_clips * wtiles_per_clip
    int hop;
    int n_mels;
    int bulk_in;             // 1: TMA bulk loads allowed (alignment checked on the host)
    int bulk_out;            // 1: TMA bulk stores allowed
    int layout;              // 0 frame-major, 1 mel-major
    int fft_size;            // for num_frames(lens[clip])
    // constant tables (global memory, staged into shared memory once per CTA)
    const float2* window;    // plan 400: [10 workers] float4 (cos th_2t, cos th_2t+1, sin th_2t, sin th_2t+1), th_c = 2 pi c/400
                             // plan 512: [32 n1][16 c] floats, the (Hann or zero-padded Povey) window itself
    const float4* twiddle;   // [10 i][10 workers]  W_400^(t*2i), W_400^(t*(2i+1)) as (re,im,re,im); row 20-t uses the
                             // conjugates + an output rotation
    const float2* rot10;     // [20]  W_40^(-c): row 10 is pre-rotated on the write side so worker 0 fits the same scheme
    const float2* proj;      // plan 512: [proj_ktot][32] (weight, __int_as_float(row))
                             // plan 400: floats, [proj_ktot4][32] weights (entry-major) then [proj_ktot4/4][32][4] (lane-major quads)
    const int* proj_meta;    // [kMaxMpl] K_s | exchange flag << 16, then [kMaxMpl][32] mel index or -1, then [kMaxMpl][32] first bin
                             // of the lane's window (the K_s consecutive power rows its entries multiply), then [kMaxMpl][32] the
                             // lane holding the other half of a split band (own lane if none)
    int proj_ktot;           // plan 400: entries rounded up to a multiple of 4
    float floor_val;         // 1e-10 (Whisper) — floor applied to the *unscaled* energy
    float log_mul;           // log10(2) (Whisper)
    int normalize;           // 1: per-frame max-8 clamp and (x+4)/4
    int frame_len;           // samples per frame before zero padding (fft_size for Whisper, 400 for Kaldi / NeMo)
    float preemph;           // Kaldi / NeMo pre-emphasis coefficient
    int frame_offset;        // NeMo: first sample of frame 0 relative to the clip (-200 when centred, +56 otherwise)
    float log_add;           // NeMo: log_zero_guard added to the energy before ln()
    int out_row_stride;      // mel-major output: floats between mel rows (NeMo: padded frame count)
    int mm_aligned8;         // mel-major output: every mel row of every tile starts on an 8-byte boundary
    int cmn_fused;           // Kaldi: CMN inside the fused kernel (one CTA per clip at a time), see melspec512_kernel
    int ps_down, ps_up;      // pair prescale: how far the floor / guard allows a frame to be scaled down / up (see pair_prescale)
    int vec_out;             // frame-major output rows are 16-byte aligned (float4 stores when TMA stores are not used)
    int n_clips;
    int smem_cmn;            // [NWARPS][128] column sums + [128] means (floats)
    int tile_order;          // plan 400: 0 = every warp owns a contiguous range of tiles, 1 = the CTA does and its warps interleave
    int smem_mmoff;          // plan 400, KSPEC 5 / 6: [3 * n_mels] ints, the mel-major store offsets of a staged tile
    int mm_sync;             // plan 400, tile_order 1: CTA barrier every mm_sync passes (0: never)
    // shared-memory carve-up (bytes from the start of dynamic smem), computed on the host
    int smem_win, smem_tw, smem_rot, smem_proj, smem_meta, smem_warp0, smem_warp_stride, smem_stage_off, smem_pcm_off, smem_scr_off;
};

constexpr int kMaxMpl = 4;
constexpr int kMetaInts = kMaxMpl + 3 * kMaxMpl * 32;

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// Non-blocking phase test (1 = the phase with this parity has completed).  The tile loops issue it late in a pass for the NEXT
// pass's PCM, so the shared-memory round trip of the mbarrier instruction (queued behind the pass's loads and stores on a pipe
// that is 80 % busy) is not on the critical path at the top of the next pass; mbar_wait() remains the fallback.
#ifndef MS_REFILL_FULL
#define MS_REFILL_FULL 1
#endif
#ifndef MS_EARLY_TEST
#define MS_EARLY_TEST 1
#endif
#ifndef MS_TMA_LEAD
#define MS_TMA_LEAD 1
#endif
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
// TMA 1-D bulk copy global -> shared, completion reported as bytes on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// TMA 1-D bulk copy shared -> global (bulk async-group completion).
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
// L2 eviction policies for the bulk copies of the plan-512 kernel.  There the PCM is loaded evict_first (it must not push the output
// rows of the clips in flight out of L2), and rows that a later pass touches again (Kaldi CMN) are written evict_last and released
// by that pass.  Plan 400 uses plain loads: measured with the hint, the 240-sample halo a warp re-reads one pass later missed L2
// (DRAM reads 657 -> 750 MB per cfg2 launch, profiles/README.md).
__device__ __forceinline__ uint64_t l2_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
#ifdef MELSPEC_NO_L2_HINTS   // A/B switch
    (void)pol;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
    return;
#endif
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g_hint(void* dst, uint32_t src, uint32_t bytes, uint64_t pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(src), "r"(bytes), "l"(pol)
                 : "memory");
}
// TMA bulk reduction: global[dst + i] += shared[src + i] for bytes / 4 floats, performed at the L2 (no data comes back to the SM).
__device__ __forceinline__ void bulk_reduce_add_f32(void* dst, uint32_t src, uint32_t bytes, uint64_t pol) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.L2::cache_hint.add.f32 [%0], [%1], %2, %3;" ::"l"(dst), "r"(src),
                 "r"(bytes), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// One elected lane of the (converged) warp arms the mbarrier with the tile's byte count and issues its chunk copies; a chunk of 0
// bytes is skipped.  A single predicated instruction sequence: no divergent region, no per-copy election loop.
__device__ __forceinline__ void tma_load_chunks4(uint32_t bar, uint32_t total, uint32_t dst0, uint32_t dst_stride, const void* src,
                                                 uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3, uint32_t src_stride) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        ".reg .b32 d;\n\t"
        ".reg .b64 s, st;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "@p mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t"
        "setp.ne.and.u32 q, %5, 0, p;\n\t"
        "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%2], [%4], %5, [%0];\n\t"
        "add.u32 d, %2, %3;\n\t"
        "add.u64 s, %4, %9;\n\t"
        "setp.ne.and.u32 q, %6, 0, p;\n\t"
        "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [d], [s], %6, [%0];\n\t"
        "add.u32 d, d, %3;\n\t"
        "add.u64 s, s, %9;\n\t"
        "setp.ne.and.u32 q, %7, 0, p;\n\t"
        "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [d], [s], %7, [%0];\n\t"
        "add.u32 d, d, %3;\n\t"
        "add.u64 s, s, %9;\n\t"
        "setp.ne.and.u32 q, %8, 0, p;\n\t"
        "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [d], [s], %8, [%0];\n\t"
        "}" ::"r"(bar),
        "r"(total), "r"(dst0), "r"(dst_stride), "l"(src), "r"(b0), "r"(b1), "r"(b2), "r"(b3), "l"((unsigned long long)src_stride)
        : "memory");
}
// Interior tile of plan 400 (every tile but a clip's last): the four chunk sizes are compile-time constants, so the byte counts are
// immediates and nothing is tested (the general form spends ~40 uniform-pipe instructions per pass on min / max / select chains).
template <uint32_t B012, uint32_t B3, uint32_t DST_STRIDE>
__device__ __forceinline__ void tma_load_chunks4_full(uint32_t bar, uint32_t dst0, const void* src) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 st;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "@p mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %3;\n\t"
        "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%1], [%2], %4, [%0];\n\t"
        "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%1+%6], [%2+%4], %4, [%0];\n\t"
        "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%1+%7], [%2+%8], %4, [%0];\n\t"
        "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%1+%9], [%2+%10], %5, [%0];\n\t"
        "}" ::"r"(bar),
        "r"(dst0), "l"(src), "n"(3 * B012 + B3), "n"(B012), "n"(B3), "n"(DST_STRIDE), "n"(2 * DST_STRIDE), "n"(2 * B012), "n"(3 * DST_STRIDE),
        "n"(3 * B012)
        : "memory");
}
// the same with an L2 eviction policy on the copies (plan 512)
__device__ __forceinline__ void tma_load_chunks4_hint(uint32_t bar, uint32_t total, uint32_t dst0, uint32_t dst_stride, const void* src,
                                                      uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3, uint32_t src_stride, uint64_t pol,
                                                      uint32_t lead16 = 0) {   // lead16 != 0: also the 16 bytes in front of src / dst0 (counted in total)
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        ".reg .b32 d;\n\t"
        ".reg .b64 s, st;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "@p mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t"
        "setp.ne.and.u32 q, %11, 0, p;\n\t"
        "sub.u32 d, %2, 16;\n\t"
        "sub.u64 s, %4, 16;\n\t"
        "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [d], [s], 16, [%0], %10;\n\t"
        "setp.ne.and.u32 q, %5, 0, p;\n\t"
        "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%2], [%4], %5, [%0], %10;\n\t"
        "add.u32 d, %2, %3;\n\t"
        "add.u64 s, %4, %9;\n\t"
        "setp.ne.and.u32 q, %6, 0, p;\n\t"
        "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [d], [s], %6, [%0], %10;\n\t"
        "add.u32 d, d, %3;\n\t"
        "add.u64 s, s, %9;\n\t"
        "setp.ne.and.u32 q, %7, 0, p;\n\t"
        "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [d], [s], %7, [%0], %10;\n\t"
        "add.u32 d, d, %3;\n\t"
        "add.u64 s, s, %9;\n\t"
        "setp.ne.and.u32 q, %8, 0, p;\n\t"
        "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [d], [s], %8, [%0], %10;\n\t"
        "}" ::"r"(bar),
        "r"(total), "r"(dst0), "r"(dst_stride), "l"(src), "r"(b0), "r"(b1), "r"(b2), "r"(b3), "l"((unsigned long long)src_stride), "l"(pol),
        "r"(lead16)
        : "memory");
}
__device__ __forceinline__ void bulk_s2g_hint_commit_if(bool pred, void* dst, uint32_t src, uint32_t bytes, uint64_t pol) {
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t"
        "@q cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%1], [%2], %3, %4;\n\t"
        "@q cp.async.bulk.commit_group;\n\t}" ::"r"((uint32_t)pred),
        "l"(dst), "r"(src), "r"(bytes), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void bulk_wait_read0_if(bool pred) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q cp.async.bulk.wait_group.read 0;\n\t}" ::"r"((uint32_t)pred) : "memory");
}
__device__ __forceinline__ void bulk_s2g_commit_if(bool pred, void* dst, uint32_t src, uint32_t bytes) {
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t"
        "@q cp.async.bulk.global.shared::cta.bulk_group [%1], [%2], %3;\n\t"
        "@q cp.async.bulk.commit_group;\n\t}" ::"r"((uint32_t)pred),
        "l"(dst), "r"(src), "r"(bytes)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// The PCM stage of a warp is single-buffered: a pass issues the shared-memory loads of all of its samples and then the TMA refill
// for the warp's next tile.  The refill does not depend on the loaded registers, so nothing in the instruction stream makes it
// wait for loads that are still queued in the LSU - and when that queue is backed up (mel-major output: hundreds of scattered
// global stores per tile and warp) the bulk copy, one L2 round trip later, overwrote samples the last loads had not read yet
// (seen as one frame of a pair computed from the next tile's samples in about 1 launch in 100 of a ragged 128-mel mel-major
// batch; racecheck does not see the async proxy).  The loads (generic proxy) and the bulk copy (async proxy) are ordered the
// way PTX prescribes: every lane's loads -> __syncwarp() -> fence.proxy.async -> cp.async.bulk.  Measured (profiles/
// r2_refill_guard.md): 0 differing launches in 600 with the fence (5 in 600 without), +0.7 % on the plan-400 kernel.
#ifndef MS_REFILL_GUARD
#define MS_REFILL_GUARD 1   // A/B switch: 0 = no ordering (the round-1 behaviour, racy)
#endif
__device__ __forceinline__ void order_loads_before_refill() {
    if (MS_REFILL_GUARD) fence_proxy_async();
}

// log2 of a normal, positive number: the energies are floored (>= 1e-10 / FLT_EPSILON / log_zero_guard) before the
// logarithm, so the denormal fix-up that __log2f() carries (FSETP + 2 predicated FMUL/FADD per call) is dead weight
__device__ __forceinline__ float lg2_normal(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float warp_max_f32(float v) {
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));  // sm_100a: CREDUX.MAX.F32
    return r;
}

// ------------------------------------------------------------------------------------------------ pair prescale
// Two neighbouring frames travel as the real (A) and imaginary (B) part of one complex transform, so the fp32 rounding
// noise of the louder frame lands in the quieter one at the louder frame's scale.  The reference transforms every frame on
// its own in f64 (src/stft.rs:89-115), and its outputs are relative to each frame's own level (per-frame clamp, src/mel.rs:645-654;
// ln of unclamped energies, src/fbank.rs:207-221).  To keep that property, frame B is multiplied by an exact power of two 2^k
// before the transform whenever the two frames' peak levels differ by more than 2^kDeadZone.  Nothing is multiplied back:
// with E' = 2^(2k) E the energies the kernel accumulates,
//     log(max(E, v))  =  log(max(E', v 2^(2k)))  -  2k log 2          (v = the energy floor; NeMo: log(E + v), v = the guard)
// so each frame carries its own (v_q, c_q) = (v 2^(2 k_q), -2 k_q log_mul) into the epilogue, where they replace the constant
// floor of the FMNMX and turn the FMUL by log_mul into an FFMA: no extra instruction there.  A frame whose samples are all zero
// (or denormal) is "silent": its partner is scaled *down* as far as v allows, which pushes the partner's rounding noise in the
// silent frame's slot far below the floor (the silent frame then sits exactly on the floor, like the reference's).
//   pk  = (biased exponent of max|A|) | (biased exponent of max|B|) << 16, identical in all lanes of the transform
constexpr int kDeadZone = 3;     // level ratios up to 2^(3+1) ride unscaled (fp32 noise stays 3 decades under the 1e-4 contract)
constexpr int kMaxShift = 45;    // |k| <= 45: 2^(2k) v stays a normal fp32 number for every floor / guard the frontends use
__device__ __forceinline__ int pack_exponents(const float ma, const float mb) {   // ma, mb >= 0
    return (__float_as_int(ma) >> 23) | ((__float_as_int(mb) >> 23) << 16);
}
__device__ __forceinline__ float scale_pow4(const float v, const int k) {          // v * 2^(2k), v and the result normal
    return __int_as_float(__float_as_int(v) + (k << 24));
}
__device__ __forceinline__ float pow2i(const int k) { return __int_as_float((127 + k) << 23); }
// returns (ka, kb) and the table entry (vA, cA, vB, cB); kdown / kup = how far v allows a frame to be scaled down / up
__device__ __forceinline__ void pair_prescale(const int pk, const float v, const float log_mul, const int kdown, const int kup,
                                              int& ka, int& kb, float4& tab) {
    const int ea = pk & 0xffff, eb = pk >> 16;
    int d = min(max(ea - eb, -kdown), kup);
    if (abs(d) <= kDeadZone) d = 0;
    ka = 0;
    kb = d;
    if (ea == 0 || eb == 0) {          // a silent frame: scale its partner down (both silent: nothing to do)
        kb = (eb != 0) ? -kdown : 0;
        ka = (ea != 0) ? -kdown : 0;
    }
    tab = make_float4(scale_pow4(v, ka), (float)(-2 * ka) * log_mul, scale_pow4(v, kb), (float)(-2 * kb) * log_mul);
}

// ------------------------------------------------------------------------------------------------ DFT codelets
// 5-point forward DFT, in place on (r[k*S], i[k*S]) k = 0..4.
#define MS_C1 0.30901699437494745f
#define MS_C2 (-0.80901699437494745f)
#define MS_S1 0.95105651629515353f
#define MS_S2 0.58778525229247314f

__device__ __forceinline__ void dft5(float& r0, float& i0, float& r1, float& i1, float& r2, float& i2, float& r3, float& i3,
                                     float& r4, float& i4) {
    const float t1r = r1 + r4, t1i = i1 + i4, t2r = r2 + r3, t2i = i2 + i3;
    const float d1r = r1 - r4, d1i = i1 - i4, d2r = r2 - r3, d2i = i2 - i3;
    const float a1r = fmaf(MS_C2, t2r, fmaf(MS_C1, t1r, r0)), a1i = fmaf(MS_C2, t2i, fmaf(MS_C1, t1i, i0));
    const float a2r = fmaf(MS_C1, t2r, fmaf(MS_C2, t1r, r0)), a2i = fmaf(MS_C1, t2i, fmaf(MS_C2, t1i, i0));
    const float b1r = fmaf(MS_S2, d2r, MS_S1 * d1r), b1i = fmaf(MS_S2, d2i, MS_S1 * d1i);
    const float b2r = fmaf(-MS_S1, d2r, MS_S2 * d1r), b2i = fmaf(-MS_S1, d2i, MS_S2 * d1i);
    r0 = r0 + t1r + t2r;
    i0 = i0 + t1i + t2i;
    // y1 = a1 - i*b1, y4 = a1 + i*b1, y2 = a2 - i*b2, y3 = a2 + i*b2     (-i*(br + i bi) = bi - i br)
    r1 = a1r + b1i; i1 = a1i - b1r;
    r4 = a1r - b1i; i4 = a1i + b1r;
    r2 = a2r + b2i; i2 = a2i - b2r;
    r3 = a2r - b2i; i3 = a2i + b2r;
}

// 20-point forward DFT, natural order in and out, Good-Thomas 4x5 (no twiddles):
//   input  n = (5a + 4b) mod 20,  output k = (5ka + 16kb) mod 20.
__device__ __forceinline__ void dft20(float (&xr)[20], float (&xi)[20]) {
    float tr[4][5], ti[4][5];
#pragma unroll
    for (int b = 0; b < 5; ++b) {
        const int n0 = (4 * b) % 20, n1 = (5 + 4 * b) % 20, n2 = (10 + 4 * b) % 20, n3 = (15 + 4 * b) % 20;
        const float s02r = xr[n0] + xr[n2], s02i = xi[n0] + xi[n2], d02r = xr[n0] - xr[n2], d02i = xi[n0] - xi[n2];
        const float s13r = xr[n1] + xr[n3], s13i = xi[n1] + xi[n3], d13r = xr[n1] - xr[n3], d13i = xi[n1] - xi[n3];
        tr[0][b] = s02r + s13r; ti[0][b] = s02i + s13i;
        tr[2][b] = s02r - s13r; ti[2][b] = s02i - s13i;
        tr[1][b] = d02r + d13i; ti[1][b] = d02i - d13r;   // d02 - i*d13
        tr[3][b] = d02r - d13i; ti[3][b] = d02i + d13r;   // d02 + i*d13
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        dft5(tr[a][0], ti[a][0], tr[a][1], ti[a][1], tr[a][2], ti[a][2], tr[a][3], ti[a][3], tr[a][4], ti[a][4]);
#pragma unroll
        for (int kb = 0; kb < 5; ++kb) {
            xr[(5 * a + 16 * kb) % 20] = tr[a][kb];
            xi[(5 * a + 16 * kb) % 20] = ti[a][kb];
        }
    }
}

// ---- packed (two transforms at once) codelets: Blackwell's FADD2 / FMUL2 / FFMA2 operate on an aligned register pair,
// so one instruction advances the same butterfly of two independent DFTs.  Each float2 below holds element n of
// transform 0 in .x and of transform 1 in .y.  Rounding is per component, identical to the scalar codelets.
typedef float2 f2;
__device__ __forceinline__ f2 add2(const f2 a, const f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 sub2(const f2 a, const f2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ f2 mul2(const f2 a, const f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 mul2c(const float c, const f2 a) { return __fmul2_rn(make_float2(c, c), a); }
__device__ __forceinline__ f2 fma2(const f2 a, const f2 b, const f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 fma2c(const float c, const f2 a, const f2 b) { return __ffma2_rn(make_float2(c, c), a, b); }

__device__ __forceinline__ void dft5x2(f2& r0, f2& i0, f2& r1, f2& i1, f2& r2, f2& i2, f2& r3, f2& i3, f2& r4, f2& i4) {
    const f2 t1r = add2(r1, r4), t1i = add2(i1, i4), t2r = add2(r2, r3), t2i = add2(i2, i3);
    const f2 d1r = sub2(r1, r4), d1i = sub2(i1, i4), d2r = sub2(r2, r3), d2i = sub2(i2, i3);
    const f2 a1r = fma2c(MS_C2, t2r, fma2c(MS_C1, t1r, r0)), a1i = fma2c(MS_C2, t2i, fma2c(MS_C1, t1i, i0));
    const f2 a2r = fma2c(MS_C1, t2r, fma2c(MS_C2, t1r, r0)), a2i = fma2c(MS_C1, t2i, fma2c(MS_C2, t1i, i0));
    const f2 b1r = fma2c(MS_S2, d2r, mul2c(MS_S1, d1r)), b1i = fma2c(MS_S2, d2i, mul2c(MS_S1, d1i));
    const f2 b2r = fma2c(-MS_S1, d2r, mul2c(MS_S2, d1r)), b2i = fma2c(-MS_S1, d2i, mul2c(MS_S2, d1i));
    r0 = add2(add2(r0, t1r), t2r);
    i0 = add2(add2(i0, t1i), t2i);
    r1 = add2(a1r, b1i); i1 = sub2(a1i, b1r);
    r4 = sub2(a1r, b1i); i4 = add2(a1i, b1r);
    r2 = add2(a2r, b2i); i2 = sub2(a2i, b2r);
    r3 = sub2(a2r, b2i); i3 = add2(a2i, b2r);
}

// Two 20-point forward DFTs at once (same Good-Thomas 4x5 index maps as dft20).
__device__ __forceinline__ void dft20x2(f2 (&xr)[20], f2 (&xi)[20]) {
    f2 tr[4][5], ti[4][5];
#pragma unroll
    for (int b = 0; b < 5; ++b) {
        const int n0 = (4 * b) % 20, n1 = (5 + 4 * b) % 20, n2 = (10 + 4 * b) % 20, n3 = (15 + 4 * b) % 20;
        const f2 s02r = add2(xr[n0], xr[n2]), s02i = add2(xi[n0], xi[n2]), d02r = sub2(xr[n0], xr[n2]), d02i = sub2(xi[n0], xi[n2]);
        const f2 s13r = add2(xr[n1], xr[n3]), s13i = add2(xi[n1], xi[n3]), d13r = sub2(xr[n1], xr[n3]), d13i = sub2(xi[n1], xi[n3]);
        tr[0][b] = add2(s02r, s13r); ti[0][b] = add2(s02i, s13i);
        tr[2][b] = sub2(s02r, s13r); ti[2][b] = sub2(s02i, s13i);
        tr[1][b] = add2(d02r, d13i); ti[1][b] = sub2(d02i, d13r);
        tr[3][b] = sub2(d02r, d13i); ti[3][b] = add2(d02i, d13r);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        dft5x2(tr[a][0], ti[a][0], tr[a][1], ti[a][1], tr[a][2], ti[a][2], tr[a][3], ti[a][3], tr[a][4], ti[a][4]);
#pragma unroll
        for (int kb = 0; kb < 5; ++kb) {
            xr[(5 * a + 16 * kb) % 20] = tr[a][kb];
            xi[(5 * a + 16 * kb) % 20] = ti[a][kb];
        }
    }
}

// The same transform with the window folded into its first layer.  For residue b the layer forms the sums and differences of
// the windowed samples (n0, n2) = (4b, 10 + 4b) and (n1, n3) = (5 + 4b, 15 + 4b):  x0 w0 +- x2 w2 = fma(+-x2, w2, x0 w0) -- one
// product and two FMAs instead of two products and two adds (a fifth of the window's multiplies disappears, and the FMA rounds once).
//   q[b] = (s02, d02, s13, d13) for the real part (frame A) or the imaginary part (frame B) of the packed input
__device__ __forceinline__ void win_first_layer(const f2 x0, const f2 x1, const f2 x2, const f2 x3, const f2 w0, const f2 w1, const f2 w2,
                                                const f2 w3, f2 (&q)[4]) {
    const f2 p0 = mul2(x0, w0), p1 = mul2(x1, w1);
    q[0] = fma2(x2, w2, p0);
    q[1] = fma2(x2, make_float2(-w2.x, -w2.y), p0);
    q[2] = fma2(x3, w3, p1);
    q[3] = fma2(x3, make_float2(-w3.x, -w3.y), p1);
}
// ... and the rest of the transform: second half of the radix-4 layer (where the real and the imaginary input first meet),
// the four 5-point transforms, natural-order outputs.
__device__ __forceinline__ void dft20x2_rest(const f2 (&qr)[5][4], const f2 (&qi)[5][4], f2 (&xr)[20], f2 (&xi)[20]) {
    f2 tr[4][5], ti[4][5];
#pragma unroll
    for (int b = 0; b < 5; ++b) {
        tr[0][b] = add2(qr[b][0], qr[b][2]); ti[0][b] = add2(qi[b][0], qi[b][2]);
        tr[2][b] = sub2(qr[b][0], qr[b][2]); ti[2][b] = sub2(qi[b][0], qi[b][2]);
        tr[1][b] = add2(qr[b][1], qi[b][3]); ti[1][b] = sub2(qi[b][1], qr[b][3]);
        tr[3][b] = sub2(qr[b][1], qi[b][3]); ti[3][b] = add2(qi[b][1], qr[b][3]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        dft5x2(tr[a][0], ti[a][0], tr[a][1], ti[a][1], tr[a][2], ti[a][2], tr[a][3], ti[a][3], tr[a][4], ti[a][4]);
#pragma unroll
        for (int kb = 0; kb < 5; ++kb) {
            xr[(5 * a + 16 * kb) % 20] = tr[a][kb];
            xi[(5 * a + 16 * kb) % 20] = ti[a][kb];
        }
    }
}

// Untangle + power of two conjugate slot pairs at once.  w1 = the packed register of slot j' = N1 - j, op = the partner values
// arranged so that (w1.x, op.x) and (w1.y, op.y) are the two (u, v) pairs; S = u + v, D = u - v per half (FADD2 with a swapped-halves
// operand, SASS .LO_HI, when op is the other register with its halves exchanged).  Powers: |A|^2 = Sr^2 + Di^2, |B|^2 = Si^2 + Dr^2.
__device__ __forceinline__ void untangle2(const f2 w1r, const f2 w1i, const f2 opr, const f2 opi, float2& plo, float2& phi) {
    const f2 sr = add2(w1r, opr), si = add2(w1i, opi), dr = sub2(w1r, opr), di = sub2(w1i, opi);
    const f2 di2 = mul2(di, di), dr2 = mul2(dr, dr);
    plo = make_float2(fmaf(sr.x, sr.x, di2.x), fmaf(si.x, si.x, dr2.x));   // scalar FMAs: the results land as (|A|^2, |B|^2) pairs
    phi = make_float2(fmaf(sr.y, sr.y, di2.y), fmaf(si.y, si.y, dr2.y));
}

// ---- power-of-two codelets for the 512-point plan -------------------------------------------------------------
// x *= W_32^k = exp(-2*pi*i*k/32), k a compile-time constant after unrolling (trivial factors cost nothing)
__device__ __forceinline__ void mul_w32(float& r, float& i, const int k) {
    constexpr float C[16] = {1.f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f, 0.70710678118654757f,
                             0.55557023301960229f, 0.38268343236508984f, 0.19509032201612833f, 0.f, -0.19509032201612819f,
                             -0.38268343236508973f, -0.55557023301960196f, -0.70710678118654746f, -0.83146961230254535f,
                             -0.92387953251128674f, -0.98078528040323043f};
    constexpr float S[16] = {0.f, 0.19509032201612825f, 0.38268343236508978f, 0.55557023301960218f, 0.70710678118654746f,
                             0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f, 1.f, 0.98078528040323043f,
                             0.92387953251128674f, 0.83146961230254546f, 0.70710678118654757f, 0.55557023301960218f,
                             0.38268343236508989f, 0.19509032201612861f};
    const int kk = k & 31;
    if (kk == 0) return;
    if (kk == 8) { const float t = r; r = i; i = -t; return; }            // -i
    if (kk == 16) { r = -r; i = -i; return; }
    if (kk == 24) { const float t = r; r = -i; i = t; return; }           // +i
    const float c = kk < 16 ? C[kk] : -C[kk - 16], sn = kk < 16 ? S[kk] : -S[kk - 16];
    const float nr = fmaf(i, sn, r * c), ni = fmaf(-r, sn, i * c);       // (r + i*im)(c - i*s)
    r = nr; i = ni;
}

__device__ __forceinline__ void dft4(float& r0, float& i0, float& r1, float& i1, float& r2, float& i2, float& r3, float& i3) {
    const float s02r = r0 + r2, s02i = i0 + i2, d02r = r0 - r2, d02i = i0 - i2;
    const float s13r = r1 + r3, s13i = i1 + i3, d13r = r1 - r3, d13i = i1 - i3;
    r0 = s02r + s13r; i0 = s02i + s13i;
    r2 = s02r - s13r; i2 = s02i - s13i;
    r1 = d02r + d13i; i1 = d02i - d13r;   // d02 - i*d13
    r3 = d02r - d13i; i3 = d02i + d13r;   // d02 + i*d13
}

// 16-point forward DFT on x[S*n], n = 0..15, natural order in and out (4 x 4 Cooley-Tukey, constant twiddles).
template <int S, int LEN>
__device__ __forceinline__ void dft16(float (&xr)[LEN], float (&xi)[LEN], const int base) {
    float tr[4][4], ti[4][4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        float r0 = xr[base + S * b], i0 = xi[base + S * b], r1 = xr[base + S * (4 + b)], i1 = xi[base + S * (4 + b)];
        float r2 = xr[base + S * (8 + b)], i2 = xi[base + S * (8 + b)], r3 = xr[base + S * (12 + b)], i3 = xi[base + S * (12 + b)];
        dft4(r0, i0, r1, i1, r2, i2, r3, i3);
        mul_w32(r1, i1, 2 * b); mul_w32(r2, i2, 4 * b); mul_w32(r3, i3, 6 * b);   // W_16^(b*ka)
        tr[0][b] = r0; ti[0][b] = i0; tr[1][b] = r1; ti[1][b] = i1; tr[2][b] = r2; ti[2][b] = i2; tr[3][b] = r3; ti[3][b] = i3;
    }
#pragma unroll
    for (int ka = 0; ka < 4; ++ka) {
        dft4(tr[ka][0], ti[ka][0], tr[ka][1], ti[ka][1], tr[ka][2], ti[ka][2], tr[ka][3], ti[ka][3]);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) { xr[base + S * (ka + 4 * kb)] = tr[ka][kb]; xi[base + S * (ka + 4 * kb)] = ti[ka][kb]; }
    }
}

// 32-point forward DFT, natural order in and out: two interleaved 16-point DFTs + one radix-2 stage.
__device__ __forceinline__ void dft32(float (&xr)[32], float (&xi)[32]) {
    dft16<2, 32>(xr, xi, 0);   // E[ka] lands in x[2*ka]
    dft16<2, 32>(xr, xi, 1);   // O[ka] lands in x[2*ka + 1]
    float er[16], ei[16], orr[16], oi[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        er[k] = xr[2 * k]; ei[k] = xi[2 * k]; orr[k] = xr[2 * k + 1]; oi[k] = xi[2 * k + 1];
        mul_w32(orr[k], oi[k], k);
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        xr[k] = er[k] + orr[k]; xi[k] = ei[k] + oi[k];
        xr[k + 16] = er[k] - orr[k]; xi[k + 16] = ei[k] - oi[k];
    }
}

// ------------------------------------------------------------------------------------------------ plan-400 constants
namespace p400 {
constexpr int N = 400;
constexpr int FPW = 6;          // frames per warp pass (3 complex FFTs)
constexpr int ZROW = 33;        // 16-byte units per Z slot row: 20 complex x 3 FFTs = 30 units + 3 pad  (33 = 1 mod 8)
constexpr int ZSLOTS = 20;
constexpr int ZBYTES = ZSLOTS * ZROW * 16;   // 10560 per warp
// power slab (reuses the Z slab): one plane per FFT, plane[bin] = float2(|A|^2, |B|^2), bins in natural order so that a mel
// band is a run of consecutive rows.  Plane origins (in float2 units) are 0, 6, 0 mod 16: with lane = 10 g + t that is
// the best a 64-bit store can do (3 wavefronts instead of 2; see tools/smem_model.py), reads are conflict-free.
constexpr int PPLANE1 = 214, PPLANE2 = 416;
constexpr int PBYTES = (PPLANE2 + 201 + 1) * 8;   // 4944
constexpr int STAGE_MAX = ZBYTES - PBYTES;   // room for the 6 x n_mels output rows behind the power rows
constexpr int CHUNK = 320;      // samples per TMA bulk copy (two hops of 160)
constexpr int PAD320 = 20;      // words of padding after each chunk in the staged PCM tile: with lane = 10 g + t the 64-bit
                                // loads of a half-warp then fall into 16 distinct bank pairs
constexpr int CS320 = CHUNK + PAD320;
constexpr int NCHUNK = 4;       // a warp tile spans 5*160 + 400 = 1200 samples = 3.75 chunks
__host__ __device__ constexpr int slot_of_row(int r) { return r <= 10 ? r : 30 - r; }
}  // namespace p400

// ------------------------------------------------------------------------------------------------ the fused kernel
// Warp-autonomous pipeline: every warp owns its own shared-memory slab (PCM stage, Z/power slab, output stage) and
// its own mbarrier, walks its own sequence of "warp tiles" (6 consecutive frames of one clip) and never meets a
// CTA-wide barrier after setup, so the warps of an SM drift into different phases and the FMA, LSU and TMA pipes
// overlap.  The PCM stage is single-buffered: a pass reads all of its samples into registers first, so the TMA load
// of the warp's next tile is issued right after that and lands during the rest of the pass.
// KSPEC selects a compile-time projection schedule: 0 = entry counts per slot read from the table (any filterbank),
// 1 = the Whisper 80-mel / fft-400 bank, whose slots hold 14, 4 and 2 entries (loops fully unrolled: no loop control,
// no register rotation, all table and power loads of a slot in flight together), 2 = the same bank with its 13 longest
// bands split over two lanes (8, 5 and 2 entries, halves added with a warp shuffle in slots 0 and 1).  The host picks it by
// comparing counts.
template <int NWARPS, int MPL, bool HOP160, int KSPEC>
__global__ void __launch_bounds__(NWARPS * 32, 1) melspec400_kernel(const KParams p) {
    using namespace p400;
    extern __shared__ __align__(128) unsigned char smem[];
    // KSPEC 3 = KSPEC 1 for the launch shape every large batch has (frame-major output, aligned buffers so that both TMA paths
    // apply, no per-clip lengths): those run-time switches become compile-time constants
    // KSPEC 4 = the same for the Slaney 128-mel bank (Whisper large-v3: 9, 4, 2 and 1 entries in its four slots)
    // KSPEC 5 = KSPEC 3 with the mel-major (`interleave_frames`, whisper.cpp) output layout instead: 8-byte aligned rows, full tiles
    // stored from a per-lane offset table (ragged tiles take the generic path below)
    // KSPEC 6 = the same for the 128-mel bank (KSPEC 4's schedule)
    constexpr bool FAST = KSPEC >= 3, FASTMM = KSPEC == 5 || KSPEC == 6;
    constexpr int KS_ = (KSPEC == 3 || KSPEC == 5) ? 1 : KSPEC == 6 ? 4 : KSPEC;
    constexpr int MM_MELS = KSPEC == 6 ? 128 : 80;
    const bool f_bulk_in = FAST ? true : (p.bulk_in != 0), f_bulk_out = FAST ? !FASTMM : (p.bulk_out != 0), f_norm = FAST ? true : (p.normalize != 0);
    const int f_layout = FAST ? (FASTMM ? 1 : 0) : p.layout;
    const int32_t* const f_lens = FAST ? nullptr : p.lens;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform: the tile-loop state lives in uniform registers
    const int lane = threadIdx.x & 31;
    const int l30 = lane < 30 ? lane : 29;   // lanes 30,31 shadow lane 29 (same addresses, same values)
    const int g = l30 / 10;                  // which of the warp's 3 FFTs: frames fw0 + 2g (re) and fw0 + 2g + 1 (im)
    const int t = l30 - 10 * g;              // worker within the FFT

    const float* s_projw = reinterpret_cast<const float*>(smem + p.smem_proj);
    const int* s_meta = reinterpret_cast<const int*>(smem + p.smem_meta);
    unsigned char* s_warp = smem + p.smem_warp0 + warp * p.smem_warp_stride;
    float4* s_z = reinterpret_cast<float4*>(s_warp);          // Z exchange slab, later reused as the power slab
    float2* s_p = reinterpret_cast<float2*>(s_warp);
    float* s_stage = reinterpret_cast<float*>(s_warp + p.smem_stage_off);
    float* s_pcm = reinterpret_cast<float*>(s_warp + p.smem_pcm_off);
    float4* s_scr = reinterpret_cast<float4*>(s_warp + p.smem_scr_off);   // pair prescale: (floor, log offset) of frames A and B per FFT
    constexpr bool LATE_LOAD = (NWARPS == 16);   // see the tile loop
    const uint32_t bar = smem_u32(smem + 8 * warp);           // this warp's "PCM landed" mbarrier
    // lanes of this FFT at ring distance 1, 2 | 3, 6, 9 (six bits each): the two-round all-reduce of the pair prescale
    const int ring = (10 * g + (t + 1) % 10) | (10 * g + (t + 2) % 10) << 6 | (10 * g + (t + 3) % 10) << 12 | (10 * g + (t + 6) % 10) << 18 |
                     (10 * g + (t + 9) % 10) << 24;

    // ---- one-time setup: tables into shared memory, barriers, per-lane window / twiddle registers
    for (int i = threadIdx.x; i < 2 * p.proj_ktot * 32; i += NWARPS * 32)
        reinterpret_cast<float*>(smem + p.smem_proj)[i] = reinterpret_cast<const float*>(p.proj)[i];
    for (int i = threadIdx.x; i < kMetaInts; i += NWARPS * 32) reinterpret_cast<int*>(smem + p.smem_meta)[i] = p.proj_meta[i];
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // window and twiddle tables are lane-dependent (10 distinct rows, shared by the warp's 3 FFTs); they are read
    // from shared memory each pass instead of pinning 78 registers, which is what lets 12 warps live on an SM
    // three copies (one per FFT of the warp), [i][g][t]: a warp's LDS.128 then covers 30 consecutive 16-byte units instead of
    // 10 units read three times in a lane order that collides inside the quarter-warps
    for (int i = threadIdx.x; i < 300; i += NWARPS * 32) reinterpret_cast<float4*>(smem + p.smem_tw)[i] = p.twiddle[10 * (i / 30) + (i % 30) % 10];
    // Periodic Hann window of columns 2t, 2t+1, evaluated on the fly (two packed FFMA2 per row instead of a table read):
    //   w[20*n1 + c] = 0.5 - 0.5 cos(2 pi n1/20 + th_c) = 0.5 + WC[n1] cos(th_c) + WS[n1] sin(th_c),  th_c = 2 pi c/400
    // the per-worker phase factors (cos th_2t, cos th_2t+1, sin th_2t, sin th_2t+1) sit in shared memory, one copy per lane:
    // with 12 warps per SM there is no register to park them in between passes (one conflict-free LDS.128 per pass)
    for (int i = threadIdx.x; i < 30; i += NWARPS * 32)
        reinterpret_cast<float4*>(smem + p.smem_win)[i] = __ldg(reinterpret_cast<const float4*>(p.window) + i % 10);
    const float4* s_wth = reinterpret_cast<const float4*>(smem + p.smem_win) + l30;
    constexpr float WC[20] = {-0.5f, -0.47552825814757677f, -0.40450849718747373f, -0.29389262614623657f, -0.15450849718747373f,
                              0.f, 0.15450849718747367f, 0.29389262614623651f, 0.40450849718747367f, 0.47552825814757677f,
                              0.5f, 0.47552825814757688f, 0.40450849718747378f, 0.29389262614623662f, 0.15450849718747378f,
                              0.f, -0.15450849718747361f, -0.29389262614623646f, -0.40450849718747367f, -0.47552825814757677f};
    constexpr float WS[20] = {0.f, 0.1545084971874737f, 0.29389262614623657f, 0.40450849718747373f, 0.47552825814757677f,
                              0.5f, 0.47552825814757682f, 0.40450849718747373f, 0.29389262614623662f, 0.15450849718747375f,
                              0.f, -0.15450849718747345f, -0.29389262614623651f, -0.40450849718747367f, -0.47552825814757677f,
                              -0.5f, -0.47552825814757682f, -0.40450849718747378f, -0.29389262614623668f, -0.15450849718747381f};
    const float4* s_tw = reinterpret_cast<const float4*>(smem + p.smem_tw) + l30;
    // the worker's 10 twiddle quads stay in registers (156 registers per thread at 12 warps, no spills, since the TMA refill is
    // issued between the sample loads and the first butterfly layer instead of after it)
#ifdef MELSPEC_TW_SMEM   // A/B switch (tools/ab_bench.sh): twiddles read from shared memory every pass
    constexpr bool TW_IN_REGS = (NWARPS <= 8);
#else
    constexpr bool TW_IN_REGS = (HOP160 && NWARPS <= 12) || (NWARPS <= 8);   // (other hops load and window on the fly, 16 warps have 128 registers)
#endif
    float4 twreg[10];
    if (TW_IN_REGS) {
#pragma unroll
        for (int i = 0; i < 10; ++i) twreg[i] = __ldg(p.twiddle + 10 * i + t);
    }
    for (int i = threadIdx.x; i < 30; i += NWARPS * 32) reinterpret_cast<float4*>(smem + p.smem_rot)[i] = reinterpret_cast<const float4*>(p.rot10)[i % 10];
    const float4* s_rot = reinterpret_cast<const float4*>(smem + p.smem_rot) + l30;   // (rx.x, rx.y, ry.x, ry.y): W_40^(-c), c = 2t, 2t+1
    // KSPEC 5 / 6: where float2 i of a staged [mels][3 x float2] tile goes, relative to the tile's first output column:
    // (i / 3) * row stride + 2 (i % 3) floats; i = lane + 32 k, one table row per k
    int* const s_mmoff = reinterpret_cast<int*>(smem + p.smem_mmoff);
    if (FASTMM)
        for (int i = threadIdx.x; i < 3 * MM_MELS; i += NWARPS * 32) s_mmoff[i] = (i / 3) * p.out_row_stride + 2 * (i % 3);
    __syncthreads();

    const int hop = HOP160 ? 160 : p.hop;
    const int need = (FPW - 1) * hop + N;    // samples a warp tile spans
    int pr_off[MPL], mel_of[MPL];
#pragma unroll
    for (int s = 0; s < MPL; ++s) { pr_off[s] = s_meta[kMaxMpl + kMaxMpl * 32 + s * 32 + lane]; mel_of[s] = s_meta[kMaxMpl + s * 32 + lane]; }

    // Stage the PCM of warp tile `wt` into this warp's buffer (TMA bulk copies issued by one lane).
    auto issue_load = [&](int clip, int tin) {
        order_loads_before_refill();   // (callers: __syncwarp() after the pass's last sample loads)
        const int fw0 = tin * FPW;
        const long long s0 = (long long)fw0 * hop;
        const long long left = (long long)p.n_samples - s0;
        const int avail = left < need ? (int)left : need;
        const float* src = p.pcm + (long long)clip * p.clip_stride + s0;
        if (FAST && MS_REFILL_FULL && left >= need) {   // (warp-uniform) interior tile: constant chunk sizes
            constexpr uint32_t NEED = (FPW - 1) * 160 + 400;   // (FAST implies HOP160, fft 400)
            tma_load_chunks4_full<CHUNK * 4u, (NEED - 3 * CHUNK) * 4u, CS320 * 4u>(bar, smem_u32(s_pcm), src);
        } else if (FAST) {   // (HOP160, aligned: the whole warp is converged here)
            const int a0 = min(CHUNK, avail), a1 = max(0, min(CHUNK, avail - CHUNK)), a2 = max(0, min(CHUNK, avail - 2 * CHUNK)),
                      a3 = max(0, min(CHUNK, avail - 3 * CHUNK));
            tma_load_chunks4(bar, (uint32_t)avail * 4u, smem_u32(s_pcm), CS320 * 4u, src, a0 * 4u, a1 * 4u, a2 * 4u, a3 * 4u, CHUNK * 4u);
        } else if (f_bulk_in) {
            if (lane == 0) {
                mbar_arrive_expect_tx(bar, (uint32_t)avail * 4u);
                if (HOP160) {
#pragma unroll
                    for (int k = 0; k < NCHUNK; ++k)
                        if (CHUNK * k < avail)
                            bulk_g2s(smem_u32(s_pcm + k * CS320), src + CHUNK * k, (uint32_t)min(CHUNK, avail - CHUNK * k) * 4u, bar);
                } else {
                    bulk_g2s(smem_u32(s_pcm), src, (uint32_t)avail * 4u, bar);
                }
            }
        } else {   // unaligned input: cooperative copy (same layout), then a plain arrive
            for (int i = lane; i < avail; i += 32) s_pcm[HOP160 ? i + PAD320 * (i / CHUNK) : i] = __ldg(src + i);
            __syncwarp();
            if (lane == 0) mbar_arrive(bar);
        }
    };

    // every warp owns a contiguous range of warp tiles (no division in the loop, neighbouring tiles share their
    // 240-sample halo through L2)
    // Mel-major output (tile_order 1): the CTA owns a contiguous range and its warps take neighbouring tiles (warp, warp + NWARPS, ...).
    // A mel row then receives the 24-byte pieces of up to twelve neighbouring tiles at about the same time, so its L2 lines fill
    // up and leave quickly; with per-warp ranges the partially written lines of all 1 776 warps sit in L2 for five passes each,
    // which pushed the PCM halo out of L2 (measured: DRAM 657 -> 802 MB read, 300 -> 406 MB written per cfg2 launch).
    const bool interleaved_tiles = (FAST && !FASTMM) ? false : (p.tile_order != 0);
    const int tstep = interleaved_tiles ? NWARPS : 1;
    int clip, tin, cnt, cnt_cta = 0;
    if (!interleaved_tiles) {
        const int nwt = gridDim.x * NWARPS, gw = blockIdx.x * NWARPS + warp;
        const int base = p.n_wtiles / nwt, rem = p.n_wtiles - base * nwt;
        const int lo = gw * base + min(gw, rem);
        cnt = base + (gw < rem ? 1 : 0);
        clip = lo / p.wtiles_per_clip;
        tin = lo - clip * p.wtiles_per_clip;
    } else {
        const int nct = gridDim.x, b = blockIdx.x;
        const int base = p.n_wtiles / nct, rem = p.n_wtiles - base * nct;
        const int lo = b * base + min(b, rem) + warp, cntb = base + (b < rem ? 1 : 0);
        cnt = cntb > warp ? (cntb - warp + NWARPS - 1) / NWARPS : 0;
        cnt_cta = (cntb + NWARPS - 1) / NWARPS;
        clip = lo / p.wtiles_per_clip;
        tin = lo - clip * p.wtiles_per_clip;
    }
    if (cnt > 0) issue_load(clip, tin);

    uint32_t pcm_ready = 0;   // (warp-uniform) result of the early phase test of the previous pass
    for (int it = 0, since_sync = 0; it < (interleaved_tiles ? cnt_cta : cnt); ++it) {
        // interleaved tile order: the warps of a CTA must stay near each other in the tile sequence, or the rows' partially written L2
        // lines pile up again (left alone they drift apart by many passes); a CTA barrier every few passes bounds the drift
        if (interleaved_tiles) {
            if (p.mm_sync > 0 && ++since_sync == p.mm_sync) { since_sync = 0; __syncthreads(); }
            if (it >= cnt) continue;   // (this warp's share is one tile shorter)
        }
        const int fw0 = tin * FPW;   // first frame of this pass
        int nfr = p.frames_per_clip;
        if (f_lens) {
            const int len = min(f_lens[clip], p.n_samples);
            nfr = len < p.fft_size ? 0 : (len - p.fft_size) / hop + 1;
        }
        const int nvalid = max(0, min(FPW, nfr - fw0));          // warp-uniform

        if (!pcm_ready) mbar_wait(bar, it & 1);
        pcm_ready = 0;

        // ------------------------------------------------------------------ step 1: window + column DFTs
        // The two column transforms of a worker (columns 2t, 2t+1) advance together in packed FADD2/FMUL2/FFMA2 instructions;
        // re = frame A, im = frame B.  QR[b] / QI[b] = the window-folded first layer of the 20-point transform (win_first_layer).
        f2 QR[5][4], QI[5][4];
        const bool va = 2 * g < nvalid, vb = 2 * g + 1 < nvalid;   // ragged tail: missing frames are exact zeros
        const float4 wth = *s_wth;
        const f2 w_cos = make_float2(wth.x, wth.y), w_sin = make_float2(wth.z, wth.w);
        auto win = [&](const int n1) { return fma2c(WC[n1], w_cos, fma2c(WS[n1], w_sin, make_float2(0.5f, 0.5f))); };
        float2 x[28];
        if (HOP160) {
            // frames A and B overlap by 240 samples: B[n1] = A[n1 + 8], so 28 loads cover both (element m is
            // sample 320g + 20m + 2t of the tile; chunk boundary at m = 16).  Conflict-free: a half-warp's 16 float2
            // addresses are 10 consecutive units of one FFT + 6 of the next, 170 = 10 (mod 16) units further on
            const float* px = s_pcm + g * CS320 + 2 * t;
#pragma unroll
            for (int m = 0; m < 28; ++m) x[m] = *reinterpret_cast<const float2*>(px + 20 * m + (m >= 16 ? PAD320 : 0));
        } else if (nvalid > 0) {
            const float* pa = s_pcm + 2 * g * hop + 2 * t;
            const float* pb = pa + hop;
            auto ld = [&](const float* q, const int n1, const bool v) {
                return v ? make_float2(q[20 * n1], q[20 * n1 + 1]) : make_float2(0.f, 0.f);
            };
#pragma unroll
            for (int b = 0; b < 5; ++b) {
                const int n0 = (4 * b) % 20, n1 = (5 + 4 * b) % 20, n2 = (10 + 4 * b) % 20, n3 = (15 + 4 * b) % 20;
                const f2 w0 = win(n0), w1 = win(n1), w2 = win(n2), w3 = win(n3);
                win_first_layer(ld(pa, n0, va), ld(pa, n1, va), ld(pa, n2, va), ld(pa, n3, va), w0, w1, w2, w3, QR[b]);
                win_first_layer(ld(pb, n0, vb), ld(pb, n1, vb), ld(pb, n2, vb), ld(pb, n3, vb), w0, w1, w2, w3, QI[b]);
            }
        }
        __syncwarp();   // every lane has issued its sample loads (HOP160: their values are consumed after the refill below, which keeps
                        // only the 56 sample registers live across the TMA issue); issue_load() orders them before the bulk copy
        const int cur_clip = clip;
        tin += tstep;
        while (tin >= p.wtiles_per_clip) { tin -= p.wtiles_per_clip; ++clip; }   // (warp-uniform)
        // 12-warp build: the refill goes out now.  16-warp build (LATE_LOAD): the PCM stage shares its first 3.7 KB with the tail
        // of the exchange slab, which is live until the row loads of step 3 are done, so the refill goes out after those.
        if ((!LATE_LOAD || nvalid == 0) && it + 1 < cnt) issue_load(clip, tin);
        if (nvalid == 0) continue;
        if (HOP160) {
#pragma unroll
            for (int b = 0; b < 5; ++b) {
                const int n0 = (4 * b) % 20, n1 = (5 + 4 * b) % 20, n2 = (10 + 4 * b) % 20, n3 = (15 + 4 * b) % 20;
                const f2 w0 = win(n0), w1 = win(n1), w2 = win(n2), w3 = win(n3);
                win_first_layer(x[n0], x[n1], x[n2], x[n3], w0, w1, w2, w3, QR[b]);
                win_first_layer(x[n0 + 8], x[n1 + 8], x[n2 + 8], x[n3 + 8], w0, w1, w2, w3, QI[b]);
            }
            if (nvalid != FPW) {   // ragged tail (warp-uniform): frames past the clip's last one are exact zeros
#pragma unroll
                for (int b = 0; b < 5; ++b)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        QR[b][i] = va ? QR[b][i] : make_float2(0.f, 0.f);
                        QI[b][i] = vb ? QI[b][i] : make_float2(0.f, 0.f);
                    }
            }
        }

        bool resc;   // warp-uniform: some pair of this pass is scaled
        {   // pair prescale (see pair_prescale).  Level of a frame = max |first-layer value| = max (|x0 w0| + |x2 w2|) over its windowed
            // sample pairs: within a factor 2 of the peak windowed sample.  All-reduced over the FFT's 10 lanes.
            float ma = 0.f, mb = 0.f;
#pragma unroll
            for (int b = 0; b < 5; ++b)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    ma = fmaxf(fmaxf(ma, fabsf(QR[b][i].x)), fabsf(QR[b][i].y));   // FMNMX3 with |.| source modifiers
                    mb = fmaxf(fmaxf(mb, fabsf(QI[b][i].x)), fabsf(QI[b][i].y));
                }
            int pk = pack_exponents(ma, mb);
            {   // max is idempotent: ring distances {0, 1, 2} then {0, 3, 6, 9} of the result cover all 10 lanes in two dependent
                // rounds (the shuffles of a round are independent), and every lane ends with the same word
                const int a1 = __shfl_sync(0xffffffffu, pk, ring), a2 = __shfl_sync(0xffffffffu, pk, ring >> 6);
                pk = __vmaxu2(__vmaxu2(pk, a1), a2);
                const int b1 = __shfl_sync(0xffffffffu, pk, ring >> 12), b2 = __shfl_sync(0xffffffffu, pk, ring >> 18),
                          b3 = __shfl_sync(0xffffffffu, pk, ring >> 24);
                pk = __vmaxu2(__vmaxu2(pk, b1), __vmaxu2(b2, b3));
            }
            int ka, kb;
            float4 tab;
            pair_prescale(pk, p.floor_val, p.log_mul, p.ps_down, p.ps_up, ka, kb, tab);
            resc = __any_sync(0xffffffffu, (ka | kb) != 0);
            if (resc) {   // rare (onsets, decays, digital silence next to sound): exact scaling; only then does the epilogue read the table
                s_scr[g] = tab;
                const float ra = pow2i(ka), rb = pow2i(kb);
#pragma unroll
                for (int b = 0; b < 5; ++b)
#pragma unroll
                    for (int i = 0; i < 4; ++i) { QR[b][i] = mul2c(ra, QR[b][i]); QI[b][i] = mul2c(rb, QI[b][i]); }
            }
        }
        f2 PR[20], PI[20];
        dft20x2_rest(QR, QI, PR, PI);
        {   // row 10 carries an extra W_40^(-c) so that worker 0 can treat it like a "row 20 - t"
            const float4 rot = *s_rot;
            const f2 rx = make_float2(rot.x, rot.y), ry = make_float2(rot.z, rot.w);
            const f2 nr = fma2(make_float2(-PI[10].x, -PI[10].y), ry, mul2(PR[10], rx));
            const f2 ni = fma2(PR[10], ry, mul2(PI[10], rx));
            PR[10] = nr; PI[10] = ni;
        }
        bulk_wait_read0_if(lane == 0);   // the previous pass's bulk store (its rows live inside this slab) is done (predicated, no branch)
        __syncwarp();
#pragma unroll
        for (int k1 = 0; k1 < 20; ++k1)     // unit = (re col 2t, re col 2t+1, im col 2t, im col 2t+1)
            s_z[ZROW * slot_of_row(k1) + l30] = make_float4(PR[k1].x, PR[k1].y, PI[k1].x, PI[k1].y);
        __syncwarp();

        // ------------------------------------------------------------------ step 3: twiddle + row DFTs
        // XR[n] = (row t, row 20-t) real parts, XI[n] = imaginary parts: again two transforms per packed instruction
        f2 XR[20], XI[20];
        {
            const float4* z1 = s_z + ZROW * t + 10 * g;          // row stride 33 = 1 (mod 8) units: a quarter-warp's 8 loads
            const float4* z2 = s_z + ZROW * (10 + t) + 10 * g;   // (t + 2g mod 8 all different) never share a bank group
#pragma unroll
            for (int i = 0; i < 10; ++i) {
                const float4 v = z1[i], u = z2[i], w = TW_IN_REGS ? twreg[i] : s_tw[30 * i];   // v,u = (re n, re m, im n, im m); w = (wr n, wi n, wr m, wi m)
                const int n = 2 * i, m = 2 * i + 1;
                XR[n] = make_float2(v.x * w.x - v.z * w.y, fmaf(u.x, w.x, u.z * w.y));         // x * tw ,  y * conj(tw)
                XI[n] = make_float2(fmaf(v.x, w.y, v.z * w.x), u.z * w.x - u.x * w.y);
                XR[m] = make_float2(v.y * w.z - v.w * w.w, fmaf(u.y, w.z, u.w * w.w));
                XI[m] = make_float2(fmaf(v.y, w.w, v.w * w.z), u.w * w.z - u.y * w.w);
            }
        }
        __syncwarp();   // every lane has its rows in registers: the slab may now be overwritten with powers
        if (LATE_LOAD && it + 1 < cnt) issue_load(clip, tin);
        dft20x2(XR, XI);   // .x = X (row t); .y = D with the row's spectrum Y[m] = D[(m + 1) % 20]
        {
            // Pair slot j: generic worker (rows a, 20-a): (X[j], Y[19-j])  -> bin a+20j (j<10) or its mirror.
            // Worker 0 (rows 0, 10): j<10: (Y[j], Y[19-j]) -> bin 10+20j;  j>=10: (X[j], X[20-j]) -> bin 20(20-j).
            // The powers go to this FFT's plane in natural bin order: slot j < 10 is bin 20j + t (worker 0: 20j + 10),
            // slot j >= 10 is the mirror bin 20(20-j) - t, so the ten workers of an FFT store ten consecutive rows.
            const bool t0 = (t == 0);
            float2* const pl = s_p + (g == 0 ? 0 : g == 1 ? PPLANE1 : PPLANE2);
            float2* const p_lo = pl + (t0 ? 10 : t);
            float2* const p_hi = pl - t;
            // Slots j and 20 - j are formed together.  Generic worker: slot 20-j = (X[20-j], D[j]), slot j = (X[j], D[20-j]), i.e.
            // register 20-j against register j with its halves exchanged.  Worker 0: slot 20-j = (X[20-j], X[j]), slot j = (D[20-j], D[j+1]).
#pragma unroll
            for (int j = 1; j < 10; ++j) {
                const f2 opr = t0 ? make_float2(XR[j].x, XR[j + 1].y) : make_float2(XR[j].y, XR[j].x);
                const f2 opi = t0 ? make_float2(XI[j].x, XI[j + 1].y) : make_float2(XI[j].y, XI[j].x);
                float2 phi, plo;
                untangle2(XR[20 - j], XI[20 - j], opr, opi, phi, plo);   // .x halves -> slot 20-j, .y halves -> slot j
                p_lo[20 * j] = plo;
                p_hi[20 * j] = phi;
            }
#pragma unroll
            for (int j = 0; j < 20; j += 10) {   // slots 0 and 10 pair a register with itself
                float ur = XR[j].x, ui = XI[j].x, vr = XR[j].y, vi = XI[j].y;
                if (j == 0) { ur = t0 ? XR[1].y : ur; ui = t0 ? XI[1].y : ui; }
                else        { vr = t0 ? XR[10].x : vr; vi = t0 ? XI[10].x : vi; }
                const float sr = ur + vr, di = ui - vi, si = ui + vi, dr = ur - vr;
                const float pwa = fmaf(sr, sr, di * di);   // 4|A[k]|^2
                const float pwb = fmaf(si, si, dr * dr);   // 4|B[k]|^2
                if (j == 0) p_lo[0] = make_float2(pwa, pwb);
                else        p_hi[200] = make_float2(pwa, pwb);
            }
        }
        __syncwarp();

        // ------------------------------------------------------------------ banded mel projection + log (+ normalise)
        // Lane l owns up to MPL mels (slot s: mel meta[s][l]).  A mel's non-zero weights are a run of consecutive bins, so
        // its entries are K_s consecutive rows of the three power planes starting at the lane's window start (immediate
        // offsets, no per-entry address); the host places the windows so that the 16 lanes of a half-warp start at 16
        // different rows mod 16 (conflict-free LDS.64).  acc[g] = (frame 2g, frame 2g+1) advances as one packed FFMA2.
        float lg[MPL][FPW];
        float mx[FPW];
#pragma unroll
        for (int q = 0; q < FPW; ++q) mx[q] = -3.0e38f;
        {
            constexpr int KS[4] = {KS_ == 4 ? 9 : KS_ == 2 ? 8 : 14, KS_ == 2 ? 5 : 4, 2, KS_ == 4 ? 1 : 0};   // KS_ != 0: the Whisper 80-mel (128-mel) / fft-400 bank
            constexpr bool EXS[4] = {KS_ == 2, KS_ == 2, false, false};         // slots whose split bands are summed by shuffle
            const float4* wq = reinterpret_cast<const float4*>(s_projw + p.proj_ktot * 32) + lane;   // [quad][lane] x 4 weights
            const float* wt = s_projw + lane;                                                          // [entry][lane]
            int eoff = 0;
            float flq[FPW], cq[FPW];   // pair prescale: (floor, log offset) of the six frames
#pragma unroll
            for (int q = 0; q < FPW; ++q) { flq[q] = p.floor_val; cq[q] = 0.f; }
            if (resc) {
                const float4 ps0 = s_scr[0], ps1 = s_scr[1], ps2 = s_scr[2];
                flq[0] = ps0.x; flq[1] = ps0.z; flq[2] = ps1.x; flq[3] = ps1.z; flq[4] = ps2.x; flq[5] = ps2.z;
                cq[0] = ps0.y; cq[1] = ps0.w; cq[2] = ps1.y; cq[3] = ps1.w; cq[4] = ps2.y; cq[5] = ps2.w;
            }
#pragma unroll
            for (int s = 0; s < MPL; ++s) {
                const float2* pr = s_p + pr_off[s];
                f2 acc0 = make_float2(0.f, 0.f), acc1 = acc0, acc2 = acc0;
                if (KS_ != 0) {
                    float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int e = 0; e < KS[s]; ++e) {
                        const int ge = eoff + e;
                        if ((ge & 3) == 0 || e == 0) w4 = wq[(ge >> 2) * 32];
                        const float w = (ge & 3) == 0 ? w4.x : (ge & 3) == 1 ? w4.y : (ge & 3) == 2 ? w4.z : w4.w;
                        const f2 ww = make_float2(w, w);
                        acc0 = fma2(ww, pr[e], acc0);
                        acc1 = fma2(ww, pr[PPLANE1 + e], acc1);
                        acc2 = fma2(ww, pr[PPLANE2 + e], acc2);
                    }
                    eoff += KS[s];
                } else {
                    const int K = s_meta[s] & 0xffff;
#pragma unroll 2
                    for (int e = 0; e < K; ++e) {
                        const float w = wt[(eoff + e) * 32];
                        const f2 ww = make_float2(w, w);
                        acc0 = fma2(ww, pr[e], acc0);
                        acc1 = fma2(ww, pr[PPLANE1 + e], acc1);
                        acc2 = fma2(ww, pr[PPLANE2 + e], acc2);
                    }
                    eoff += K;
                }
                if (KS_ != 0 ? EXS[s] : (s_meta[s] >> 16) != 0) {   // warp-uniform: add the other half of split bands
                    const int pl = s_meta[kMaxMpl + 2 * kMaxMpl * 32 + s * 32 + lane];
                    const float sel = pl != lane ? 1.0f : 0.0f;
                    const f2 o0 = make_float2(__shfl_sync(0xffffffffu, acc0.x, pl), __shfl_sync(0xffffffffu, acc0.y, pl));
                    const f2 o1 = make_float2(__shfl_sync(0xffffffffu, acc1.x, pl), __shfl_sync(0xffffffffu, acc1.y, pl));
                    const f2 o2 = make_float2(__shfl_sync(0xffffffffu, acc2.x, pl), __shfl_sync(0xffffffffu, acc2.y, pl));
                    const f2 ss = make_float2(sel, sel);
                    acc0 = fma2(ss, o0, acc0); acc1 = fma2(ss, o1, acc1); acc2 = fma2(ss, o2, acc2);
                }
                const float a[FPW] = {acc0.x, acc0.y, acc1.x, acc1.y, acc2.x, acc2.y};
#pragma unroll
                for (int q = 0; q < FPW; ++q) {
                    lg[s][q] = fmaf(p.log_mul, lg2_normal(fmaxf(a[q], flq[q])), cq[q]);
                    mx[q] = fmaxf(mx[q], lg[s][q]);
                }
            }
        }
        if (MS_EARLY_TEST && it + 1 < cnt) pcm_ready = mbar_test(bar, (it + 1) & 1);   // (the refill went out at the top of this pass)
        if (f_norm) {
#pragma unroll
            for (int q = 0; q < FPW; ++q) mx[q] = warp_max_f32(mx[q]) - 8.0f;
        }

        // ------------------------------------------------------------------ store
        if (f_layout == 0) {
            // the output rows are staged in the slab right behind the power rows (both dead once the next pass's
            // Z exchange starts; the wait above orders the bulk store against that)
#pragma unroll
            for (int s = 0; s < MPL; ++s) {
                const int mel = mel_of[s];
                float v[FPW];
#pragma unroll
                for (int q = 0; q < FPW; ++q) v[q] = f_norm ? fmaf(fmaxf(lg[s][q], mx[q]), 0.25f, 1.0f) : lg[s][q];
                // six predicated stores (lanes without a mel in this slot skip them) instead of a divergent region per slot
                const uint32_t a0 = smem_u32(s_stage + max(mel, 0)), rs = (uint32_t)p.n_mels * 4u;
                asm volatile(
                    "{\n\t.reg .pred q;\n\t.reg .b32 a;\n\t"
                    "setp.ge.s32 q, %0, 0;\n\t"
                    "@q st.shared.f32 [%1], %3;\n\t"
                    "add.u32 a, %1, %2;\n\t@q st.shared.f32 [a], %4;\n\t"
                    "add.u32 a, a, %2;\n\t@q st.shared.f32 [a], %5;\n\t"
                    "add.u32 a, a, %2;\n\t@q st.shared.f32 [a], %6;\n\t"
                    "add.u32 a, a, %2;\n\t@q st.shared.f32 [a], %7;\n\t"
                    "add.u32 a, a, %2;\n\t@q st.shared.f32 [a], %8;\n\t}" ::"r"(mel),
                    "r"(a0), "r"(rs), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5])
                    : "memory");
            }
            float* dst = p.out + (long long)cur_clip * p.out_clip_stride + (long long)fw0 * p.n_mels;
            const int nout = nvalid * p.n_mels;
            if (f_bulk_out) {
                fence_proxy_async();
                __syncwarp();
                bulk_s2g_commit_if(lane == 0, dst, smem_u32(s_stage), (uint32_t)nout * 4u);   // (predicated, no branch)
            } else {
                __syncwarp();
                for (int i = lane; i < nout; i += 32) dst[i] = s_stage[i];
                __syncwarp();
            }
        } else if (nvalid == FPW) {
            // mel-major / interleave_frames layout, full tile: a mel row receives 6 consecutive floats (24 bytes).  Stage the
            // tile as [mel][3 x float2]; every lane then stores whole rows with the widest stores the row's address allows
            // (3 x STG.64, or 4 + 8 + 8 + 4 bytes when the row starts on an odd word: odd frame counts / row strides), instead of
            // 18 instructions that each scatter 4 bytes into 32 different rows.
            float2* st2 = reinterpret_cast<float2*>(s_stage);
#pragma unroll
            for (int s = 0; s < MPL; ++s) {
                const int mel = mel_of[s];
                if (mel >= 0) {
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        const float v0 = f_norm ? fmaf(fmaxf(lg[s][2 * u], mx[2 * u]), 0.25f, 1.0f) : lg[s][2 * u];
                        const float v1 = f_norm ? fmaf(fmaxf(lg[s][2 * u + 1], mx[2 * u + 1]), 0.25f, 1.0f) : lg[s][2 * u + 1];
                        st2[3 * mel + u] = make_float2(v0, v1);
                    }
                }
            }
            __syncwarp();
            float* dst = p.out + (long long)cur_clip * p.out_clip_stride + fw0;
            if (FASTMM) {   // aligned rows: 3 float2 per mel, offsets from the table (80 mels: 7 full rounds + 16 lanes; 128 mels: 12 rounds)
#pragma unroll
                for (int k = 0; k < (3 * MM_MELS + 31) / 32; ++k)
                    if (32 * k + 32 <= 3 * MM_MELS || lane < 3 * MM_MELS - 32 * k) *reinterpret_cast<float2*>(dst + s_mmoff[lane + 32 * k]) = st2[lane + 32 * k];
            } else
            if (p.mm_aligned8) {   // all rows 8-byte aligned: walk the staged tile linearly, three lanes per row, one STG.64 each
                int row = lane / 3, u = lane - 3 * row;
                for (int i = lane; i < 3 * p.n_mels; i += 32) {
                    *reinterpret_cast<float2*>(dst + (long long)row * p.out_row_stride + 2 * u) = st2[i];
                    row += 10; u += 2;
                    if (u >= 3) { u -= 3; ++row; }
                }
            } else
            for (int row = lane; row < p.n_mels; row += 32) {
                const float2 a = st2[3 * row], b = st2[3 * row + 1], c = st2[3 * row + 2];
                float* r = dst + (long long)row * p.out_row_stride;
                if ((reinterpret_cast<uintptr_t>(r) & 7) == 0) {
                    reinterpret_cast<float2*>(r)[0] = a; reinterpret_cast<float2*>(r)[1] = b; reinterpret_cast<float2*>(r)[2] = c;
                } else {
                    r[0] = a.x;
                    *reinterpret_cast<float2*>(r + 1) = make_float2(a.y, b.x);
                    *reinterpret_cast<float2*>(r + 3) = make_float2(b.y, c.x);
                    r[5] = c.y;
                }
            }
            __syncwarp();
        } else {   // mel-major: out[clip][mel][frame], ragged tile or unaligned rows
            float* dst = p.out + (long long)cur_clip * p.out_clip_stride + fw0;
#pragma unroll
            for (int s = 0; s < MPL; ++s) {
                const int mel = mel_of[s];
                if (mel >= 0) {
#pragma unroll
                    for (int q = 0; q < FPW; ++q) {
                        const float v = f_norm ? fmaf(fmaxf(lg[s][q], mx[q]), 0.25f, 1.0f) : lg[s][q];
                        if (q < nvalid) dst[(long long)mel * p.out_row_stride + q] = v;
                    }
                }
            }
            __syncwarp();
        }
    }
    if (lane == 0) bulk_wait0();   // all bulk stores of this warp have landed before the CTA retires
}

// ---- packed power-of-two codelets (two transforms per instruction, see the f2 helpers above) ----------------------
__device__ __forceinline__ void mul_w32x2(f2& r, f2& i, const int k) {
    constexpr float C[16] = {1.f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f, 0.70710678118654757f,
                             0.55557023301960229f, 0.38268343236508984f, 0.19509032201612833f, 0.f, -0.19509032201612819f,
                             -0.38268343236508973f, -0.55557023301960196f, -0.70710678118654746f, -0.83146961230254535f,
                             -0.92387953251128674f, -0.98078528040323043f};
    constexpr float S[16] = {0.f, 0.19509032201612825f, 0.38268343236508978f, 0.55557023301960218f, 0.70710678118654746f,
                             0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f, 1.f, 0.98078528040323043f,
                             0.92387953251128674f, 0.83146961230254546f, 0.70710678118654757f, 0.55557023301960218f,
                             0.38268343236508989f, 0.19509032201612861f};
    const int kk = k & 31;
    if (kk == 0) return;
    if (kk == 8) { const f2 t = r; r = i; i = make_float2(-t.x, -t.y); return; }
    if (kk == 16) { r = make_float2(-r.x, -r.y); i = make_float2(-i.x, -i.y); return; }
    if (kk == 24) { const f2 t = r; r = make_float2(-i.x, -i.y); i = t; return; }
    const float c = kk < 16 ? C[kk] : -C[kk - 16], sn = kk < 16 ? S[kk] : -S[kk - 16];
    const f2 nr = fma2c(sn, i, mul2c(c, r)), ni = fma2c(-sn, r, mul2c(c, i));
    r = nr; i = ni;
}

__device__ __forceinline__ void dft4x2(f2& r0, f2& i0, f2& r1, f2& i1, f2& r2, f2& i2, f2& r3, f2& i3) {
    const f2 s02r = add2(r0, r2), s02i = add2(i0, i2), d02r = sub2(r0, r2), d02i = sub2(i0, i2);
    const f2 s13r = add2(r1, r3), s13i = add2(i1, i3), d13r = sub2(r1, r3), d13i = sub2(i1, i3);
    r0 = add2(s02r, s13r); i0 = add2(s02i, s13i);
    r2 = sub2(s02r, s13r); i2 = sub2(s02i, s13i);
    r1 = add2(d02r, d13i); i1 = sub2(d02i, d13r);
    r3 = sub2(d02r, d13i); i3 = add2(d02i, d13r);
}

// Two 16-point forward DFTs at once, natural order in and out.
__device__ __forceinline__ void dft16x2(f2 (&xr)[16], f2 (&xi)[16]) {
    f2 tr[4][4], ti[4][4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        f2 r0 = xr[b], i0 = xi[b], r1 = xr[4 + b], i1 = xi[4 + b], r2 = xr[8 + b], i2 = xi[8 + b], r3 = xr[12 + b], i3 = xi[12 + b];
        dft4x2(r0, i0, r1, i1, r2, i2, r3, i3);
        mul_w32x2(r1, i1, 2 * b); mul_w32x2(r2, i2, 4 * b); mul_w32x2(r3, i3, 6 * b);
        tr[0][b] = r0; ti[0][b] = i0; tr[1][b] = r1; ti[1][b] = i1; tr[2][b] = r2; ti[2][b] = i2; tr[3][b] = r3; ti[3][b] = i3;
    }
#pragma unroll
    for (int ka = 0; ka < 4; ++ka) {
        dft4x2(tr[ka][0], ti[ka][0], tr[ka][1], ti[ka][1], tr[ka][2], ti[ka][2], tr[ka][3], ti[ka][3]);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) { xr[ka + 4 * kb] = tr[ka][kb]; xi[ka + 4 * kb] = ti[ka][kb]; }
    }
}

// 32-point forward DFT of one sequence: its even and odd halves run as the two lanes of a packed 16-point DFT.
//   in : er[a] = (x[2a], x[2a+1]) real parts, ei[a] imaginary parts;  out: xr/xi natural order
__device__ __forceinline__ void dft32_packed(f2 (&er)[16], f2 (&ei)[16], float (&xr)[32], float (&xi)[32]) {
    dft16x2(er, ei);   // .x = E[k], .y = O[k]
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        float orr = er[k].y, oi = ei[k].y;
        mul_w32(orr, oi, k);
        xr[k] = er[k].x + orr; xi[k] = ei[k].x + oi;
        xr[k + 16] = er[k].x - orr; xi[k + 16] = ei[k].x - oi;
    }
}

// ================================================================================================ plan 512
// N = 512 = 32 x 16: a warp owns 2 complex FFTs = 4 frames per pass, 16 lanes per FFT.
//   step 1  lane (c, g) = (lane & 15, lane >> 4) transforms column c (elements x[16*n1 + c]) with a 32-point DFT
//   step 3  lane (t, g) = (lane >> 1, lane & 1) owns rows t and 32-t (t = 0: rows 0 and 16): two 16-point DFTs
// Same conjugate-twiddle / output-rotation / pre-rotated middle row scheme as plan 400.  KALDI adds the fbank.rs prologue
// (per-frame DC removal, pre-emphasis with look-back, Povey window, 400 samples zero-padded to 512) and drops the Whisper
// normalisation (natural log of the floored energies; CMN is a second small kernel).
namespace p512 {
constexpr int N = 512;
constexpr int FPW = 4;
constexpr int ZROWB = 144;                   // bytes per Z row: 16 complex + 16 B pad  (9 units: odd => conflict-free LDS.128)
constexpr int ZSLABB = 32 * ZROWB + 64;      // 4672 B per FFT (292 units = 4 mod 8: the two FFTs of a quarter-warp never collide)
constexpr int ZBYTES = 2 * ZSLABB;           // 9344 per warp
constexpr int kCmnRows = 24;                 // rows per bulk reduction of the fused CMN (24 x 80 floats = 7680 bytes)
constexpr int SCRBYTES = 32;                 // behind the slab: the pair prescale's per-frame (floor, log offset), one float4 per FFT
constexpr int PBYTES = 258 * 16;             // power rows in natural bin order 0..256, one float4 (A0, B0, A1, B1) per row
constexpr int STAGE_MAX = ZBYTES - PBYTES;
constexpr int CHUNK = 320;
constexpr int PAD = 16;
constexpr int CS = CHUNK + PAD;
constexpr int NCHUNK = 4;                    // 3*160 + 512 = 992 samples
__host__ __device__ constexpr int slot_of_row(int r) { return r <= 16 ? r : 48 - r; }
// TMA_LEAD (Kaldi look-back sample): the 16 bytes in front of a warp's PCM stage must be free.  The stage starts at the next multiple
// of 128 after the slab and the prescale scratch (launch_device: smem_pcm_off), which leaves this much unused in front of it:
static_assert((ZBYTES + SCRBYTES + 127) / 128 * 128 - (ZBYTES + SCRBYTES) >= 16, "no room for the look-back sample in front of the PCM stage");
}  // namespace p512

// MODE 0: Whisper fft 512.  MODE 1: Kaldi fbank.  MODE 2: NeMo BatchLogMel (whole-waveform pre-emphasis, frames may
// hang over both ends of the clip: the missing samples are zero-filled in the stage, reference src/mel.rs:344-348,685-706).
// MODE 3: the same with per-clip lengths (ragged batch), its own instantiation so that the dense mode pays nothing for it.
#ifndef TWREG512
#define TWREG512 false   // A/B switch: twiddles of the row transforms in registers (measured 3 % slower for the Whisper mode)
#endif
// FAST: the launch shape every large dense batch has (aligned buffers so that the TMA paths apply, no per-clip lengths, the
// frontend's own layout) compiled in, like KSPEC 3 of melspec400_kernel.
// KSCHED: compile-time projection schedule (entries per slot), like KSPEC of melspec400_kernel: 0 = counts read from the table (any
// filterbank), 1 = Slaney 80-mel without the Nyquist bin (Whisper fft 512: 18, 5, 2), 2 = the Kaldi 80-bin bank (16, 6, 2), 3 = Slaney
// 80-mel with the Nyquist bin (NeMo: 19, 5, 2), 4 = Slaney 128-mel (NeMo 128 / Whisper large-v3 style at fft 512: 12, 6, 3, 2).  The
// loops are then fully unrolled (no loop control, weights read as LDS.128 quads, all loads of a slot in flight together); the host
// picks the schedule by comparing the counts of the table it built.
__host__ __device__ constexpr int ksched512(int k, int s) {
    constexpr int T[5][4] = {{0, 0, 0, 0}, {18, 5, 2, 0}, {16, 6, 2, 0}, {19, 5, 2, 0}, {12, 6, 3, 2}};
    return T[k][s];
}
// MM (with FAST, MODE 0): the mel-major / interleave_frames layout compiled in instead of the frame-major one.
template <int NWARPS, int MPL, int MODE, bool FAST = false, int KSCHED = 0, bool MM = false>
__global__ void __launch_bounds__(NWARPS * 32, 1) melspec512_kernel(const KParams p) {
    using namespace p512;
    extern __shared__ __align__(128) unsigned char smem[];
    const bool f_bulk_in = FAST ? true : (p.bulk_in != 0), f_norm = FAST ? (MODE == 0) : (p.normalize != 0);
    const int f_layout = FAST ? ((MODE >= 2 || MM) ? 1 : 0) : p.layout;
    const bool f_bulk_out = FAST ? (MODE < 2 && !MM) : (p.bulk_out != 0);
    const int32_t* const f_lens = FAST ? nullptr : p.lens;
    constexpr bool KALDI = MODE == 1, NEMO = MODE == 2 || MODE == 3, RAGGED = MODE == 3, FRAME400 = MODE != 0;
    // Where the pre-emphasis look-back sample comes from (see the tile loop).  Kaldi only: the NeMo modes measured 0.3 % (80 mel) to 12 %
    // (128 mel) slower with it (profiles/r2_early_test_tma_lead.md)
    constexpr bool TMA_LEAD = MS_TMA_LEAD && FAST && KALDI;
    constexpr int NLOAD = FRAME400 ? 35 : 42;  // rows of 16 samples covering frames A and B (B = A shifted by 10 rows)
    constexpr int NROW = FRAME400 ? 25 : 32;   // non-zero rows of a frame (400 samples zero-padded to 512)

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform: the tile-loop state lives in uniform registers
    const int lane = threadIdx.x & 31;
    const int c = lane & 15, g1 = lane >> 4;   // step-1 role
    const int t = lane >> 1, g3 = lane & 1;    // step-3 role

    const float* s_projw = reinterpret_cast<const float*>(smem + p.smem_proj);
    const int* s_meta = reinterpret_cast<const int*>(smem + p.smem_meta);
    unsigned char* s_warp = smem + p.smem_warp0 + warp * p.smem_warp_stride;
    float4* s_p4 = reinterpret_cast<float4*>(s_warp);
    float2* s_p2 = reinterpret_cast<float2*>(s_warp);
    float* s_stage = reinterpret_cast<float*>(s_warp + p.smem_stage_off);
    float* s_pcm = reinterpret_cast<float*>(s_warp + p.smem_pcm_off);
    float4* s_scr = reinterpret_cast<float4*>(s_warp + p.smem_scr_off);   // pair prescale: (floor or guard, log offset) of frames A and B per FFT
    const uint32_t bar = smem_u32(smem + 8 * warp);
    const uint64_t pol_in = l2_evict_first();                 // PCM is read once: it must not displace output rows in L2

    // tables: window [32][16] floats, twiddles [8 i][16 t] float4 = (W_512^(t*2i), W_512^(t*(2i+1)))
    for (int i = threadIdx.x; i < 512; i += NWARPS * 32) reinterpret_cast<float*>(smem + p.smem_win)[i] = reinterpret_cast<const float*>(p.window)[i];
    for (int i = threadIdx.x; i < 128; i += NWARPS * 32) reinterpret_cast<float4*>(smem + p.smem_tw)[i] = p.twiddle[i];
    for (int i = threadIdx.x; i < (KSCHED != 0 ? 2 : 1) * p.proj_ktot * 32; i += NWARPS * 32)   // weights, [entry][lane] (+ [quad][lane][4])
        reinterpret_cast<float*>(smem + p.smem_proj)[i] = reinterpret_cast<const float*>(p.proj)[i];
    for (int i = threadIdx.x; i < kMetaInts; i += NWARPS * 32) reinterpret_cast<int*>(smem + p.smem_meta)[i] = p.proj_meta[i];
    if (lane == 0) {
        mbar_init(bar, 1);
        if (warp == 0) {   // fused CMN, mode 2: "all warps have finished clip" x 2 parities, "sums consumed" x 2 parities
            mbar_init(smem_u32(smem + 8 * NWARPS), NWARPS * 32);         // every thread arrives with its own sums (its own release)
            mbar_init(smem_u32(smem + 8 * (NWARPS + 1)), NWARPS * 32);
            mbar_init(smem_u32(smem + 8 * (NWARPS + 2)), 32);                // the 32 threads of warp 0
            mbar_init(smem_u32(smem + 8 * (NWARPS + 3)), 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const float* s_win = reinterpret_cast<const float*>(smem + p.smem_win) + c;
    const float4* s_tw = reinterpret_cast<const float4*>(smem + p.smem_tw) + t;
    const float2 r16 = __ldg(p.rot10 + c);     // W_32^(-c): pre-rotation of row 16
#ifdef MELSPEC_TW_SMEM
    constexpr bool TW_IN_REGS = false;
#else
    constexpr bool TW_IN_REGS = TWREG512;
#endif
    float4 twreg[8];
    if (TW_IN_REGS) {
#pragma unroll
        for (int i = 0; i < 8; ++i) twreg[i] = __ldg(p.twiddle + 16 * i + t);
    }
    __syncthreads();

    const int need = (FPW - 1) * 160 + p.frame_len;
    // A/B switches (measured on B200, profiles/r2_ab_ksched512.txt): window starts and mel indices in registers: 1 - 2 % faster in
    // every mode; prescale table touched only in passes that scale something: another 1 % for Whisper-512, nothing (registers) for
    // the Kaldi and NeMo modes
#ifndef MS512_META_REGS
#define MS512_META_REGS 1
#endif
#ifndef MS512_RESC
#define MS512_RESC 1
#endif
    constexpr bool META_REGS = MS512_META_REGS && KSCHED != 0;   // window start | mel << 16 per slot, kept in registers
    constexpr bool RESC = MS512_RESC && KSCHED != 0 && MODE == 0;   // prescale table written / read only in passes that scale something
    int meta_r[MPL];
    if (META_REGS) {
#pragma unroll
        for (int s = 0; s < MPL; ++s)
            meta_r[s] = (s_meta[kMaxMpl + kMaxMpl * 32 + s * 32 + lane] & 0xffff) | (int)((unsigned)s_meta[kMaxMpl + s * 32 + lane] << 16);
    }
    auto mel_of = [&](int s) -> int { return META_REGS ? (meta_r[s] >> 16) : s_meta[kMaxMpl + s * 32 + lane]; };
    auto win_of = [&](int s) -> int { return META_REGS ? (meta_r[s] & 0xffff) : s_meta[kMaxMpl + kMaxMpl * 32 + s * 32 + lane]; };

    // NeMo, ragged batch (per-clip lengths): every clip is its own waveform zero-padded to the common width, with its own frame
    // count (src/mel.rs:387-395); samples past its length read as zeros, columns past its frame count are written as zeros
    auto clip_len = [&](int clip) -> int { return RAGGED ? max(0, min(p.lens[clip], p.n_samples)) : p.n_samples; };   // (RAGGED is never FAST)
    auto issue_load = [&](int clip, int tin) {
        order_loads_before_refill();   // (callers: __syncwarp() after the pass's last sample loads)
        const int fw0 = tin * FPW;
        const long long s0 = (long long)fw0 * 160;
        const long long left = (long long)p.n_samples - s0;
        const int avail = left < need ? (int)left : need;
        const float* src = p.pcm + (long long)clip * p.clip_stride + s0;
        if (NEMO) {
            // tile-relative range [lo, hi) that exists in the clip; everything else of [0, need) is zero (centre padding,
            // frames hanging over the end)
            const long long t0 = s0 + p.frame_offset;
            const int lo = t0 < 0 ? (int)(-t0) : 0;
            const long long endl = (long long)clip_len(clip) - t0;
            const int hi = endl < need ? (endl < lo ? lo : (int)endl) : need;
            const float* tsrc = p.pcm + (long long)clip * p.clip_stride + t0;
            if (lo > 0 || hi < need) {   // warp-uniform
                for (int i = lane; i < need; i += 32)
                    if (i < lo || i >= hi) s_pcm[i + PAD * (i / CHUNK)] = 0.f;
                __syncwarp();   // the zeros are read by other lanes; the mbarrier below only orders the TMA bytes
            }
            if (FAST && lo == 0 && hi == need) {   // interior tile: one elected lane, one predicated sequence (warp converged)
                tma_load_chunks4_hint(bar, (uint32_t)need * 4u, smem_u32(s_pcm), CS * 4u, tsrc, min(CHUNK, need) * 4u,
                                      max(0, min(CHUNK, need - CHUNK)) * 4u, max(0, min(CHUNK, need - 2 * CHUNK)) * 4u,
                                      max(0, min(CHUNK, need - 3 * CHUNK)) * 4u, CHUNK * 4u, pol_in);
            } else if (f_bulk_in && hi > lo && (!RAGGED || (hi & 3) == 0)) {   // (a ragged clip's last tile ends on any sample: bulk copies move whole 16-byte units)
                if (lane == 0) {
                    mbar_arrive_expect_tx(bar, (uint32_t)(hi - lo) * 4u);
#pragma unroll
                    for (int k = 0; k < NCHUNK; ++k) {
                        const int a = max(lo, CHUNK * k), b = min(hi, CHUNK * (k + 1));
                        if (a < b) bulk_g2s_hint(smem_u32(s_pcm + k * CS + (a - CHUNK * k)), tsrc + a, (uint32_t)(b - a) * 4u, bar, pol_in);
                    }
                }
            } else {
                for (int i = lo + lane; i < hi; i += 32) s_pcm[i + PAD * (i / CHUNK)] = __ldg(tsrc + i);
                __syncwarp();
                if (lane == 0) mbar_arrive(bar);
            }
        } else if (FAST) {   // the whole warp is converged here: one elected lane, one predicated sequence
            const int a0 = min(CHUNK, avail), a1 = max(0, min(CHUNK, avail - CHUNK)), a2 = max(0, min(CHUNK, avail - 2 * CHUNK)),
                      a3 = max(0, min(CHUNK, avail - 3 * CHUNK));
            const uint32_t lead16 = TMA_LEAD && s0 > 0 ? 16u : 0u;   // the pre-emphasis look-back sample rides in front of the tile
            tma_load_chunks4_hint(bar, (uint32_t)avail * 4u + lead16, smem_u32(s_pcm), CS * 4u, src, a0 * 4u, a1 * 4u, a2 * 4u, a3 * 4u, CHUNK * 4u, pol_in, lead16);
        } else if (f_bulk_in) {
            if (lane == 0) {
                mbar_arrive_expect_tx(bar, (uint32_t)avail * 4u);
#pragma unroll
                for (int k = 0; k < NCHUNK; ++k)
                    if (CHUNK * k < avail)
                        bulk_g2s_hint(smem_u32(s_pcm + k * CS), src + CHUNK * k, (uint32_t)min(CHUNK, avail - CHUNK * k) * 4u, bar, pol_in);
            }
        } else {
            for (int i = lane; i < avail; i += 32) s_pcm[i + PAD * (i / CHUNK)] = __ldg(src + i);
            __syncwarp();
            if (lane == 0) mbar_arrive(bar);
        }
    };

    // Tile order.  Default: warp-strided over all tiles of the launch.  Fused CMN (Kaldi, cmn_fused): the CTA works on one
    // clip at a time (clips blockIdx.x, + gridDim.x, ...; tiles warp, + NWARPS, ... inside the clip), so that when the clip
    // is finished its rows are still in L2: the CTA then subtracts the per-mel mean in place (src/fbank.rs:226-233) instead
    // of a second kernel making two more trips to HBM.  Column sums: per lane in registers (fixed order), per warp in
    // shared memory, across warps in a fixed order -> deterministic.
    const bool fused = KALDI && p.cmn_fused;
    const int wstride = gridDim.x * NWARPS;
    int wt = fused ? blockIdx.x * p.wtiles_per_clip + warp : blockIdx.x * NWARPS + warp;
    // (clip, tile in clip) of the current tile, advanced without divisions in the loop
    const int wq = wstride / p.wtiles_per_clip, wr = wstride - wq * p.wtiles_per_clip;
    int clip_f = fused ? (int)blockIdx.x : wt / p.wtiles_per_clip;
    int tile_in_clip = fused ? warp : wt - clip_f * p.wtiles_per_clip;
    auto next_tile = [&](int cur, int& tin, int& cl) -> int { // advance; returns the next global tile id
        if (!fused) {
            cl += wq; tin += wr;
            if (tin >= p.wtiles_per_clip) { tin -= p.wtiles_per_clip; ++cl; }
            return cur + wstride;
        }
        if (tin + NWARPS < p.wtiles_per_clip) { tin += NWARPS; return cur + NWARPS; }
        tin = warp;
        cl += (int)gridDim.x;
        return cl * p.wtiles_per_clip + warp;
    };
    float csum[MPL];
    int cmn_iter = 0;   // clips this CTA has finished (fused CMN, mode 2)
#pragma unroll
    for (int s = 0; s < MPL; ++s) csum[s] = 0.f;
    if (wt < p.n_wtiles) issue_load(clip_f, tile_in_clip);
    uint32_t pcm_ready = 0;   // (warp-uniform) result of the early phase test of the previous pass

    for (int it = 0; wt < p.n_wtiles; ++it) {
        const int clip = clip_f;
        const int fw0 = tile_in_clip * FPW;
        int nfr = p.frames_per_clip;
        const int len_c = clip_len(clip);
        if (!NEMO && f_lens) {
            const int len = min(f_lens[clip], p.n_samples);
            nfr = len < p.frame_len ? 0 : (len - p.frame_len) / 160 + 1;
        }
        if (RAGGED)   // centred (frame_offset < 0): len / hop + 1 frames; otherwise whole n_fft windows only
            nfr = min(nfr, len_c <= 0 ? 0 : p.frame_offset < 0 ? len_c / 160 + 1 : (len_c < N ? 0 : (len_c - N) / 160 + 1));
        const int nvalid = max(0, min(FPW, nfr - fw0));

        // pre-emphasis look-back: the sample just before the tile (only the lane that owns tile sample 0 needs it).  TMA_LEAD: it
        // arrives with the tile (issue_load puts it into the word in front of the stage); otherwise a global load, whose latency the
        // whole warp waits out at the first use of its register (2 % of the Kaldi kernel's stall samples)
        float lead = 0.f;
        const bool owns_first = FRAME400 && g1 == 0 && c == 0;   // (lane 0)
        const long long tile0 = (long long)fw0 * 160 + (NEMO ? p.frame_offset : 0);   // clip index of tile sample 0
        if (!TMA_LEAD && owns_first && tile0 > 0 && tile0 - 1 < len_c) lead = __ldg(p.pcm + (long long)clip * p.clip_stride + tile0 - 1);

        if (!pcm_ready) mbar_wait(bar, it & 1);
        pcm_ready = 0;
        if (TMA_LEAD && owns_first && tile0 > 0) lead = s_pcm[-1];

        // ------------------------------------------------------------------ step 1
        // column c: re = frame A (fw0 + 2g), im = frame B (fw0 + 2g + 1); er[a] = (re[2a], re[2a+1]) feeds the packed codelet
        f2 er[16], ei[16];
        if (nvalid > 0) {
            const bool va = 2 * g1 < nvalid, vb = 2 * g1 + 1 < nvalid;
            const float* px = s_pcm + g1 * CS + c;
            float x[NLOAD];
#pragma unroll
            for (int m = 0; m < NLOAD; ++m) x[m] = px[16 * m + PAD * (m / 20)];
            if (!FRAME400) {
#pragma unroll
                for (int a = 0; a < 16; ++a) {
                    const float w0 = s_win[32 * a], w1 = s_win[32 * a + 16];
                    er[a] = make_float2(x[2 * a] * w0, x[2 * a + 1] * w1);
                    ei[a] = make_float2(x[2 * a + 10] * w0, x[2 * a + 11] * w1);
                }
            } else {
                // d[m] = x[m] - preemph * x[m-1]; the previous sample is one word back (one chunk pad further back at a
                // chunk start); frame sums for the DC removal are reduced over the 16 lanes of the FFT
                float sa = 0.f, sb = 0.f, smid = 0.f;
                const float x0 = x[0];
#pragma unroll
                for (int m = 0; m < NLOAD; ++m) {
                    float xp;
                    if (m % 20 == 0) {
                        const int back = (c == 0) ? 1 + PAD : 1;
                        xp = (m == 0 && owns_first) ? lead : px[16 * m + PAD * (m / 20) - back];
                    } else {
                        xp = px[16 * m + PAD * (m / 20) - 1];
                    }
                    // frame sums for the DC removal: rows 10..24 belong to both frames and are summed once
                    if (KALDI && m < 10) sa += x[m];
                    if (KALDI && m >= 10 && m < 25) smid += x[m];
                    if (KALDI && m >= 25) sb += x[m];
                    x[m] = fmaf(-p.preemph, xp, x[m]);
                }
                float ka = 0.f, kb = 0.f;
                if (KALDI) {
                    sa += smid; sb += smid;
#pragma unroll
                    for (int o = 8; o >= 1; o >>= 1) {
                        sa += __shfl_xor_sync(0xffffffffu, sa, o);
                        sb += __shfl_xor_sync(0xffffffffu, sb, o);
                    }
                    const float mu_a = sa * (1.0f / 400.0f), mu_b = sb * (1.0f / 400.0f);
                    if (owns_first && fw0 == 0) x[0] = fmaf(-p.preemph, mu_a, x0);   // first frame of the clip: no look-back
                    ka = (1.0f - p.preemph) * mu_a; kb = (1.0f - p.preemph) * mu_b;
                } else {
                    // NeMo pre-emphasises the waveform before padding: the first padding sample after the clip stays zero
                    // (it would otherwise pick up -c * x[len-1]); every other padded position is 0 - c*0 already
                    const long long rel = (long long)len_c - tile0 - 320 * g1 - c;   // tile-relative index of sample `len`
                    if (rel >= 0 && rel < 16 * NLOAD && (rel & 15) == 0) {
#pragma unroll
                        for (int m = 0; m < NLOAD; ++m)
                            if (rel == 16 * m) x[m] = 0.f;
                    }
                    (void)x0;
                }
#pragma unroll
                for (int a = 0; a < 16; ++a) {
                    float r0 = 0.f, r1 = 0.f, i0 = 0.f, i1 = 0.f;
                    if (2 * a < NROW) {
                        const float w0 = s_win[32 * a];
                        r0 = (x[2 * a] - ka) * w0;
                        i0 = (x[2 * a + 10] - kb) * w0;
                    }
                    if (2 * a + 1 < NROW) {
                        const float w1 = s_win[32 * a + 16];
                        r1 = (x[2 * a + 1] - ka) * w1;
                        i1 = (x[2 * a + 11] - kb) * w1;
                    }
                    er[a] = make_float2(r0, r1);
                    ei[a] = make_float2(i0, i1);
                }
            }
            if (nvalid != FPW) {   // ragged tail (warp-uniform): frames past the clip's last one are exact zeros
#pragma unroll
                for (int a = 0; a < 16; ++a) {
                    er[a] = va ? er[a] : make_float2(0.f, 0.f);
                    ei[a] = vb ? ei[a] : make_float2(0.f, 0.f);
                }
            }
        }
        __syncwarp();
        int tin_next = tile_in_clip, clip_next = clip_f;
        const int wt_next = next_tile(wt, tin_next, clip_next);
        if (wt_next < p.n_wtiles) issue_load(clip_next, tin_next);
        if (nvalid != 0) {   // (tiles past a short clip's last frame do no work but still take part in the clip's CMN step)

        bool resc;   // warp-uniform: some frame of this pass is scaled
        {   // pair prescale (see pair_prescale): peak levels of the two prepared frames, all-reduced over the FFT's 16 lanes
            float ma = 0.f, mb = 0.f;
#pragma unroll
            for (int a = 0; a < (NROW + 1) / 2; ++a) {
                ma = fmaxf(fmaxf(ma, fabsf(er[a].x)), fabsf(er[a].y));
                mb = fmaxf(fmaxf(mb, fabsf(ei[a].x)), fabsf(ei[a].y));
            }
            int pk = pack_exponents(ma, mb);
            {   // two dependent rounds instead of four: lanes at xor distance {1, 2, 3}, then {4, 8, 12} of the result
                const int a1 = __shfl_xor_sync(0xffffffffu, pk, 1), a2 = __shfl_xor_sync(0xffffffffu, pk, 2), a3 = __shfl_xor_sync(0xffffffffu, pk, 3);
                pk = __vmaxu2(__vmaxu2(pk, a1), __vmaxu2(a2, a3));
                const int b1 = __shfl_xor_sync(0xffffffffu, pk, 4), b2 = __shfl_xor_sync(0xffffffffu, pk, 8), b3 = __shfl_xor_sync(0xffffffffu, pk, 12);
                pk = __vmaxu2(__vmaxu2(pk, b1), __vmaxu2(b2, b3));
            }
            int ka, kb;
            float4 tab;
            pair_prescale(pk, NEMO ? p.log_add : p.floor_val, p.log_mul, p.ps_down, p.ps_up, ka, kb, tab);
            resc = __any_sync(0xffffffffu, (ka | kb) != 0);
            if (!RESC || resc) s_scr[g1] = tab;
            if (resc) {   // rare: exact power-of-two scaling of a frame
                const float ra = pow2i(ka), rb = pow2i(kb);
#pragma unroll
                for (int a = 0; a < 16; ++a) { er[a] = mul2c(ra, er[a]); ei[a] = mul2c(rb, ei[a]); }
            }
        }
        float ar[32], ai[32];
        dft32_packed(er, ei, ar, ai);
        {   // row 16 carries an extra W_32^(-c)
            const float r = ar[16] * r16.x - ai[16] * r16.y, i = fmaf(ar[16], r16.y, ai[16] * r16.x);
            ar[16] = r; ai[16] = i;
        }
        bulk_wait_read0_if(lane == 0);
        __syncwarp();
        {
            unsigned char* zw = s_warp + g1 * ZSLABB + 8 * c;
#pragma unroll
            for (int k1 = 0; k1 < 32; ++k1)
                *reinterpret_cast<float2*>(zw + ZROWB * slot_of_row(k1)) = make_float2(ar[k1], ai[k1]);
        }
        __syncwarp();

        // ------------------------------------------------------------------ step 3
        f2 XR[16], XI[16];   // .x = row t, .y = row 32-t (conjugate twiddles; spectrum rotated by one)
        {
            const float4* z1 = reinterpret_cast<const float4*>(s_warp + g3 * ZSLABB + ZROWB * t);
            const float4* z2 = reinterpret_cast<const float4*>(s_warp + g3 * ZSLABB + ZROWB * (16 + t));
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 v = z1[i], u = z2[i], w = TW_IN_REGS ? twreg[i] : s_tw[16 * i];
                const int n = 2 * i, m = 2 * i + 1;
                XR[n] = make_float2(v.x * w.x - v.y * w.y, fmaf(u.x, w.x, u.y * w.y));
                XI[n] = make_float2(fmaf(v.x, w.y, v.y * w.x), u.y * w.x - u.x * w.y);
                XR[m] = make_float2(v.z * w.z - v.w * w.w, fmaf(u.z, w.z, u.w * w.w));
                XI[m] = make_float2(fmaf(v.z, w.w, v.w * w.z), u.w * w.z - u.z * w.w);
            }
        }
        __syncwarp();
        dft16x2(XR, XI);   // .x = X; .y = D with the row's spectrum Y[m] = D[(m + 1) % 16]
        {
            // Powers go to natural bin order: slot j < 8 is bin 32j + t (worker 0: 32j + 16), slot j >= 8 is the mirror bin
            // 32(16-j) - t; row = float4 (A0, B0, A1, B1), this lane fills the float2 of its FFT g3.  The 32 lanes of a
            // store cover 32 consecutive float2 (conflict-free).
            const bool t0 = (t == 0);
            float2* const p_lo = s_p2 + 2 * (t0 ? 16 : t) + g3;
            float2* const p_hi = s_p2 - 2 * t + g3;
            // slots j and 16 - j together (see melspec400_kernel): register 16-j against register j with its halves exchanged
#pragma unroll
            for (int j = 1; j < 8; ++j) {
                const f2 opr = t0 ? make_float2(XR[j].x, XR[j + 1].y) : make_float2(XR[j].y, XR[j].x);
                const f2 opi = t0 ? make_float2(XI[j].x, XI[j + 1].y) : make_float2(XI[j].y, XI[j].x);
                float2 phi, plo;
                untangle2(XR[16 - j], XI[16 - j], opr, opi, phi, plo);   // .x halves -> slot 16-j, .y halves -> slot j
                p_lo[64 * j] = plo;
                p_hi[64 * j] = phi;
            }
#pragma unroll
            for (int j = 0; j < 16; j += 8) {   // slots 0 and 8 pair a register with itself
                float ur = XR[j].x, ui = XI[j].x, vr = XR[j].y, vi = XI[j].y;
                if (j == 0) { ur = t0 ? XR[1].y : ur; ui = t0 ? XI[1].y : ui; }
                else        { vr = t0 ? XR[8].x : vr; vi = t0 ? XI[8].x : vi; }
                const float sr = ur + vr, di = ui - vi, si = ui + vi, dr = ur - vr;
                const float2 pw = make_float2(fmaf(sr, sr, di * di), fmaf(si, si, dr * dr));
                if (j == 0) p_lo[0] = pw;
                else        p_hi[64 * 8] = pw;
            }
        }
        __syncwarp();

        // ------------------------------------------------------------------ projection + log (+ normalise)
        float lg[MPL][FPW];
        float mx[FPW];
#pragma unroll
        for (int q = 0; q < FPW; ++q) mx[q] = -3.0e38f;
        {
            // windowed projection (see melspec400_kernel): the lane's K_s entries are consecutive power rows from its window
            // start, so the loads do not depend on the table and pipeline freely; weights come from [entry][lane]
            const float* wt = s_projw + lane;
            // pair prescale: (floor or guard, log offset) of the four frames
            const float v0 = NEMO ? p.log_add : p.floor_val;
            float vq[FPW] = {v0, v0, v0, v0}, cq[FPW] = {0.f, 0.f, 0.f, 0.f};
            if (!RESC || resc) {
                const float4 ps0 = s_scr[0], ps1 = s_scr[1];
                vq[0] = ps0.x; vq[1] = ps0.z; vq[2] = ps1.x; vq[3] = ps1.z;
                cq[0] = ps0.y; cq[1] = ps0.w; cq[2] = ps1.y; cq[3] = ps1.w;
            }
            const float4* wq = reinterpret_cast<const float4*>(s_projw + p.proj_ktot * 32) + lane;   // [quad][lane] x 4 weights
            int eoff = 0;
#pragma unroll
            for (int s = 0; s < MPL; ++s) {
                const float4* pr = s_p4 + win_of(s);
                f2 acc01 = make_float2(0.f, 0.f), acc23 = acc01;
                if (KSCHED != 0) {
                    float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int e = 0; e < ksched512(KSCHED, s); ++e) {
                        const int ge = eoff + e;
                        if ((ge & 3) == 0 || e == 0) w4 = wq[(ge >> 2) * 32];
                        const float w = (ge & 3) == 0 ? w4.x : (ge & 3) == 1 ? w4.y : (ge & 3) == 2 ? w4.z : w4.w;
                        const float4 pw = pr[e];
                        const f2 ww = make_float2(w, w);
                        acc01 = fma2(ww, make_float2(pw.x, pw.y), acc01);
                        acc23 = fma2(ww, make_float2(pw.z, pw.w), acc23);
                    }
                    eoff += ksched512(KSCHED, s);
                } else {
                    const int K = s_meta[s] & 0xffff;
#pragma unroll 4
                    for (int e = 0; e < K; ++e) {
                        const float w = wt[e * 32];
                        const float4 pw = pr[e];
                        const f2 ww = make_float2(w, w);
                        acc01 = fma2(ww, make_float2(pw.x, pw.y), acc01);
                        acc23 = fma2(ww, make_float2(pw.z, pw.w), acc23);
                    }
                    wt += K * 32;
                }
                const float acc[FPW] = {acc01.x, acc01.y, acc23.x, acc23.y};
#pragma unroll
                for (int q = 0; q < FPW; ++q) {
                    const float e = NEMO ? acc[q] + vq[q] : fmaxf(acc[q], vq[q]);   // ln(E + guard) / log(max(E, floor)), prescaled
                    lg[s][q] = fmaf(p.log_mul, lg2_normal(e), cq[q]);
                    mx[q] = fmaxf(mx[q], lg[s][q]);
                }
            }
        }
        if (MS_EARLY_TEST && FRAME400 && wt_next < p.n_wtiles) pcm_ready = mbar_test(bar, (it + 1) & 1);   // (the refill went out at the top of this pass; Whisper-512 measured 0.3 % slower with it)
        if (f_norm) {
#pragma unroll
            for (int q = 0; q < FPW; ++q) mx[q] = warp_max_f32(mx[q]) - 8.0f;
        }
        if (f_layout == 0) {
#pragma unroll
            for (int s = 0; s < MPL; ++s) {
                const int mel = mel_of(s);
                float v[FPW];
#pragma unroll
                for (int q = 0; q < FPW; ++q) v[q] = f_norm ? fmaf(fmaxf(lg[s][q], mx[q]), 0.25f, 1.0f) : lg[s][q];
                // four predicated stores (lanes without a mel in this slot skip them) instead of a divergent region per slot
                const uint32_t a0 = smem_u32(s_stage + max(mel, 0)), rs = (uint32_t)p.n_mels * 4u;
                asm volatile(
                    "{\n\t.reg .pred q;\n\t.reg .b32 a;\n\t"
                    "setp.ge.s32 q, %0, 0;\n\t"
                    "@q st.shared.f32 [%1], %3;\n\t"
                    "add.u32 a, %1, %2;\n\t@q st.shared.f32 [a], %4;\n\t"
                    "add.u32 a, a, %2;\n\t@q st.shared.f32 [a], %5;\n\t"
                    "add.u32 a, a, %2;\n\t@q st.shared.f32 [a], %6;\n\t}" ::"r"(mel),
                    "r"(a0), "r"(rs), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3])
                    : "memory");
            }
            float* dst = p.out + (long long)clip * p.out_clip_stride + (long long)fw0 * p.n_mels;
            const int nout = nvalid * p.n_mels;
            if (f_bulk_out) {
                fence_proxy_async();
                __syncwarp();
                if (fused) bulk_s2g_hint_commit_if(lane == 0, dst, smem_u32(s_stage), (uint32_t)nout * 4u, l2_evict_last());   // the CMN reduction comes back for these rows
                else bulk_s2g_commit_if(lane == 0, dst, smem_u32(s_stage), (uint32_t)nout * 4u);
            } else if (p.vec_out) {
                __syncwarp();
                for (int i = lane; i < nout / 4; i += 32) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(s_stage)[i];
                __syncwarp();
            } else {
                __syncwarp();
                for (int i = lane; i < nout; i += 32) dst[i] = s_stage[i];
                __syncwarp();
            }
            if (fused) {
#pragma unroll
                for (int s = 0; s < MPL; ++s)
#pragma unroll
                    for (int q = 0; q < FPW; ++q)
                        if (q < nvalid) csum[s] += lg[s][q];
            }
        } else if (nvalid == FPW) {
            // mel-major / feature-major layout, full tile: a mel row receives 4 consecutive floats (16 bytes): stage the
            // tile as [mel] x float4 and store whole rows with the widest stores the row's address allows (odd frame counts
            // make the rows start on odd words: 4 + 8 + 4 bytes)
            float4* st4 = reinterpret_cast<float4*>(s_stage);
#pragma unroll
            for (int s = 0; s < MPL; ++s) {
                const int mel = mel_of(s);
                if (mel >= 0) {
                    float v[FPW];
#pragma unroll
                    for (int q = 0; q < FPW; ++q) v[q] = f_norm ? fmaf(fmaxf(lg[s][q], mx[q]), 0.25f, 1.0f) : lg[s][q];
                    st4[mel] = make_float4(v[0], v[1], v[2], v[3]);
                }
            }
            __syncwarp();
            float* dst = p.out + (long long)clip * p.out_clip_stride + fw0;
            for (int mel = lane; mel < p.n_mels; mel += 32) {
                const float4 v = st4[mel];
                float* r = dst + (long long)mel * p.out_row_stride;
                const unsigned al = (unsigned)(reinterpret_cast<uintptr_t>(r) & 15);
                if (al == 0) {
                    *reinterpret_cast<float4*>(r) = v;
                } else if ((al & 7) == 0) {
                    *reinterpret_cast<float2*>(r) = make_float2(v.x, v.y);
                    *reinterpret_cast<float2*>(r + 2) = make_float2(v.z, v.w);
                } else {
                    r[0] = v.x;
                    *reinterpret_cast<float2*>(r + 1) = make_float2(v.y, v.z);
                    r[3] = v.w;
                }
            }
            __syncwarp();
        } else {
            float* dst = p.out + (long long)clip * p.out_clip_stride + fw0;
#pragma unroll
            for (int s = 0; s < MPL; ++s) {
                const int mel = mel_of(s);
                if (mel >= 0) {
#pragma unroll
                    for (int q = 0; q < FPW; ++q)
                        if (q < nvalid)
                            dst[(long long)mel * p.out_row_stride + q] = f_norm ? fmaf(fmaxf(lg[s][q], mx[q]), 0.25f, 1.0f) : lg[s][q];
                }
            }
            __syncwarp();
        }
        }   // nvalid != 0
        if (RAGGED && nvalid < FPW) {   // columns past a short clip's own frame count are zeros, like the pad_to columns
            const int q1 = min(FPW, p.frames_per_clip - fw0);
            float* dst = p.out + (long long)clip * p.out_clip_stride + fw0;
            for (int mel = lane; mel < p.n_mels; mel += 32)
                for (int q = nvalid; q < q1; ++q) dst[(long long)mel * p.out_row_stride + q] = 0.f;
        }

        // ------------------------------------------------------------------ fused CMN: end of this warp's share of the clip
        // Mode 2 (default): no CTA-wide barrier and no second pass through the SM.  Every warp delivers its column sums and
        // signals an mbarrier once its own bulk stores of the clip have completed, then carries on with the next clip.  Warp 0
        // alone waits for the twelve arrivals, forms the means in a fixed order and hands the subtraction to the TMA engine:
        // bulk reductions  out[row][mel] += -mean[mel]  (cp.reduce.async.bulk .add.f32, kCmnRows rows per operation) that the L2
        // applies to the rows the CTA wrote a few microseconds earlier (written evict_last, released evict_first by the
        // reduction).  x + (-m) rounds like x - m, so the result equals the two-pass form bit for bit.
        if (fused && p.cmn_fused == 2 && clip_next != clip) {
            const int par = cmn_iter & 1;
            float* s_cs = reinterpret_cast<float*>(smem + p.smem_cmn) + par * (NWARPS * 128);   // [2][NWARPS][128]
            float* s_neg = reinterpret_cast<float*>(smem + p.smem_cmn) + 2 * NWARPS * 128;      // [kCmnRows][n_mels]: -mean, replicated
            const uint32_t bar_done = smem_u32(smem + 8 * (NWARPS + par)), bar_free = smem_u32(smem + 8 * (NWARPS + 2 + par));
            if (cmn_iter >= 2) mbar_wait(bar_free, ((cmn_iter >> 1) - 1) & 1);   // warp 0 has consumed this buffer's previous sums
#pragma unroll
            for (int s = 0; s < MPL; ++s) {
                const int mel = mel_of(s);
                if (mel >= 0) s_cs[warp * 128 + mel] = csum[s];
                csum[s] = 0.f;
            }
            if (lane == 0) bulk_wait0();   // this warp's rows of the clip have landed
            mbar_arrive(bar_done);         // every thread: release of its own column sums (lane 0: after its stores)
            if (warp == 0) {
                mbar_wait(bar_done, (cmn_iter >> 1) & 1);
                if (lane == 0) bulk_wait_read0();   // the previous clip's reductions have read s_neg
                __syncwarp();
                if (nfr > 0) {
                    for (int mel = lane; mel < p.n_mels; mel += 32) {
                        float sum = 0.f;
#pragma unroll
                        for (int w = 0; w < NWARPS; ++w) sum += s_cs[w * 128 + mel];
                        const float neg = -(sum / (float)nfr);
                        for (int r = 0; r < kCmnRows; ++r) s_neg[r * p.n_mels + mel] = neg;
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        float* base = p.out + (long long)clip * p.out_clip_stride;
                        const uint64_t pol = l2_evict_first();
                        for (int r0 = 0; r0 < nfr; r0 += kCmnRows) {
                            const int rows = min(kCmnRows, nfr - r0);
                            bulk_reduce_add_f32(base + (long long)r0 * p.n_mels, smem_u32(s_neg), (uint32_t)(rows * p.n_mels) * 4u, pol);
                        }
                        bulk_commit();   // (s_neg is rewritten one clip later; every pass waits for the reads of all groups first)
                    }
                }
                mbar_arrive(bar_free);     // every thread of warp 0: its reads of the sums are done
            }
            ++cmn_iter;
        }
        // Mode 1 (MELSPEC_CMN_FUSED=1, the round-1 form, kept for A/B): block barrier, then all threads subtract in place.
        if (fused && p.cmn_fused == 1 && clip_next != clip) {
            float* s_cs = reinterpret_cast<float*>(smem + p.smem_cmn);       // [NWARPS][128]
            float* s_mean = s_cs + NWARPS * 128;                             // [128]
#pragma unroll
            for (int s = 0; s < MPL; ++s) {
                const int mel = mel_of(s);
                if (mel >= 0) s_cs[warp * 128 + mel] = csum[s];
                csum[s] = 0.f;
            }
            __syncthreads();   // every warp of the CTA finishes the same clip here (wtiles_per_clip >= NWARPS, host-checked)
            if ((int)threadIdx.x < p.n_mels) {
                float sum = 0.f;
#pragma unroll
                for (int w = 0; w < NWARPS; ++w) sum += s_cs[w * 128 + threadIdx.x];
                s_mean[threadIdx.x] = nfr > 0 ? sum / (float)nfr : 0.f;
            }
            __syncthreads();
            // the clip's nfr x n_mels floats were written by this CTA's own (plain) stores a few microseconds ago: L2 hits
            float* base = p.out + (long long)clip * p.out_clip_stride;
            const int quads = p.n_mels / 4, total = nfr * quads;
            const float4* mean4 = reinterpret_cast<const float4*>(s_mean);
            // four independent L2 loads in flight per thread: the pass is latency-bound otherwise (one CTA per SM)
            constexpr int NT = NWARPS * 32, UN = 4;
            const int step1 = NT % quads, stepu = (NT * UN) % quads;   // column of element i + NT / i + NT*UN without a division
            int mq = (int)threadIdx.x % quads;
            for (int i0 = threadIdx.x; i0 < total; i0 += NT * UN) {
                float4 v[UN];
#pragma unroll
                for (int u = 0; u < UN; ++u)
                    if (i0 + u * NT < total) v[u] = __ldcg(reinterpret_cast<const float4*>(base) + i0 + u * NT);
                int mc = mq;
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    const int i = i0 + u * NT;
                    if (i < total) {
                        const float4 m = mean4[mc];
                        v[u].x -= m.x; v[u].y -= m.y; v[u].z -= m.z; v[u].w -= m.w;
                        reinterpret_cast<float4*>(base)[i] = v[u];
                    }
                    mc += step1;
                    if (mc >= quads) mc -= quads;
                }
                mq += stepu;
                if (mq >= quads) mq -= quads;
            }
        }
        wt = wt_next;
        tile_in_clip = tin_next;
        clip_f = clip_next;
    }
    if (lane == 0) bulk_wait0();
}

// Cepstral mean normalisation of the Kaldi path (reference src/fbank.rs:226-233): out[clip][f][m] -= mean_f out[clip][f][m].
// One CTA per clip.  Vector path (n_mels % 4 == 0, 16-byte aligned rows): thread = (row group, column quad), float4
// loads with 4 independent accumulators; the column sums are reduced in a fixed order (deterministic).
__global__ void __launch_bounds__(512, 2) melspec_cmn_kernel(float* out, long long out_clip_stride, int frames_per_clip, int n_mels,
                                                             const int32_t* lens, int n_samples, int frame_len, int hop) {
    __shared__ float4 part4[512];
    __shared__ float mean_s[128];
    const int clip = blockIdx.x;
    int nfr = frames_per_clip;
    if (lens) {
        const int len = min(lens[clip], n_samples);
        nfr = len < frame_len ? 0 : (len - frame_len) / hop + 1;
    }
    if (nfr <= 0) return;
    float* base = out + (long long)clip * out_clip_stride;
    const bool vec = (n_mels % 4 == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
    if (vec) {
        const int quads = n_mels / 4;                 // <= 32
        const int rows = 512 / quads;                 // row groups working in parallel
        const int q = threadIdx.x % quads, r = threadIdx.x / quads;
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
        if (r < rows) {
            const float4* src = reinterpret_cast<const float4*>(base) + q;
            int f = r;
            for (; f + 3 * rows < nfr; f += 4 * rows) {
                const float4 v0 = src[(long long)f * quads], v1 = src[(long long)(f + rows) * quads];
                const float4 v2 = src[(long long)(f + 2 * rows) * quads], v3 = src[(long long)(f + 3 * rows) * quads];
                a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
                a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
                a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w;
                a3.x += v3.x; a3.y += v3.y; a3.z += v3.z; a3.w += v3.w;
            }
            for (; f < nfr; f += rows) {
                const float4 v0 = src[(long long)f * quads];
                a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
            }
        }
        part4[threadIdx.x] = make_float4((a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y),
                                         (a0.z + a1.z) + (a2.z + a3.z), (a0.w + a1.w) + (a2.w + a3.w));
        __syncthreads();
        if (threadIdx.x < quads) {
            float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int rr = 0; rr < rows; ++rr) {
                const float4 v = part4[rr * quads + threadIdx.x];
                sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
            }
            const float inv = 1.0f / (float)nfr;
            mean_s[4 * threadIdx.x] = sum.x * inv; mean_s[4 * threadIdx.x + 1] = sum.y * inv;
            mean_s[4 * threadIdx.x + 2] = sum.z * inv; mean_s[4 * threadIdx.x + 3] = sum.w * inv;
        }
        __syncthreads();
        if (r < rows) {
            const float4 mu = make_float4(mean_s[4 * q], mean_s[4 * q + 1], mean_s[4 * q + 2], mean_s[4 * q + 3]);
            float4* dst = reinterpret_cast<float4*>(base) + q;
            for (int f = r; f < nfr; f += rows) {
                float4 v = dst[(long long)f * quads];
                v.x -= mu.x; v.y -= mu.y; v.z -= mu.z; v.w -= mu.w;
                dst[(long long)f * quads] = v;
            }
        }
    } else {
        float* part = reinterpret_cast<float*>(part4);   // [4][128]
        const int m = threadIdx.x & 127, grp = threadIdx.x >> 7;
        float acc = 0.f;
        if (m < n_mels)
            for (int f = grp; f < nfr; f += 4) acc += base[(long long)f * n_mels + m];
        part[grp * 128 + m] = acc;
        __syncthreads();
        const float mean = (part[m] + part[128 + m] + part[256 + m] + part[384 + m]) / (float)nfr;
        if (m < n_mels)
            for (int f = grp; f < nfr; f += 4) base[(long long)f * n_mels + m] -= mean;
    }
}

// Per-feature normalisation of the NeMo frontend (reference src/mel.rs:721-749): for every mel row of a clip,
// x <- (x - mean) / (sqrt(sum (x - mean)^2 / max(F - 1, 1)) + 1e-5) over the F valid frames.  One warp per row
// (mel-major rows are contiguous), fixed-order lane-strided sums + xor-shuffle tree: deterministic.
__global__ void __launch_bounds__(256) melspec_featnorm_kernel(float* out, long long out_clip_stride, int row_stride, int frames,
                                                               int n_mels, int n_rows_total, const int32_t* lens, int n_samples,
                                                               int hop, int n_fft, int center) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n_rows_total || frames <= 0) return;
    const int clip = row / n_mels, mel = row - clip * n_mels;
    if (lens) {   // ragged batch: statistics over the clip's own valid frames (src/mel.rs:387-395, 721-749)
        const int len = min(lens[clip], n_samples);
        const int nfr = len <= 0 ? 0 : center ? len / hop + 1 : (len < n_fft ? 0 : (len - n_fft) / hop + 1);
        frames = min(frames, nfr);
        if (frames <= 0) return;
    }
    float* r = out + (long long)clip * out_clip_stride + (long long)mel * row_stride;
    if (frames <= 1024) {   // the row fits the warp's registers (32 values per lane): one read and one write, same summation order
        float v[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = lane + 32 * k < frames ? __ldcs(r + lane + 32 * k) : 0.f;
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) sum += v[k];   // (the masked tail adds exact zeros)
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum / (float)frames;
        float var = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k)
            if (lane + 32 * k < frames) { const float d = v[k] - mean; var = fmaf(d, d, var); }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
        const float inv = 1.0f / (sqrtf(var / fmaxf((float)frames - 1.0f, 1.0f)) + 1e-5f);
#pragma unroll
        for (int k = 0; k < 32; ++k)
            if (lane + 32 * k < frames) r[lane + 32 * k] = (v[k] - mean) * inv;
        return;
    }
    float sum = 0.f;
    for (int f = lane; f < frames; f += 32) sum += r[f];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)frames;
    float var = 0.f;
    for (int f = lane; f < frames; f += 32) { const float d = r[f] - mean; var = fmaf(d, d, var); }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    const float inv = 1.0f / (sqrtf(var / fmaxf((float)frames - 1.0f, 1.0f)) + 1e-5f);
    for (int f = lane; f < frames; f += 32) r[f] = (r[f] - mean) * inv;
}

// Zero padding columns [c0, c1) of every row of a batch of row-major images (interleave_frames' zero frame / min_width padding,
// src/mel.rs:497-516; NeMo's pad_to columns, src/mel.rs:336).  One thread per (row, column): a 2-D memset of 8-byte rows costs more
// than the fused kernel itself.
__global__ void __launch_bounds__(256) melspec_zero_cols_kernel(float* img, long long img_stride, int rows, long long row_stride, int c0, int c1,
                                                                long long n_imgs) {
    const int ncol = c1 - c0;
    const long long total = n_imgs * rows * ncol;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const long long r = i / ncol;
        const int cidx = (int)(i - r * ncol);
        const long long im = r / rows;
        const int row = (int)(r - im * rows);
        img[im * img_stride + (long long)row * row_stride + c0 + cidx] = 0.f;
    }
}

// ================================================================================================ 16-bit PCM input
// x / 32768 for the int16 host entry (melspec_compute_host_i16): exact in f32, so the features equal those of the f32 entry
// on the converted samples.  HBM-bound byte work: 2 bytes read + 4 written per sample; 8 samples per thread where the rows
// allow 16-byte loads.  grid = (blocks, rows).
__global__ void __launch_bounds__(256) melspec_i16_to_f32_kernel(const int16_t* in, long long in_stride, int n, float* out,
                                                                 long long out_stride, int vec) {
    const int16_t* src = in + (long long)blockIdx.y * in_stride;
    float* dst = out + (long long)blockIdx.y * out_stride;
    const float sc = 1.0f / 32768.0f;
    const int n8 = vec ? n / 8 : 0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n8; i += gridDim.x * 256) {
        const int4 v = __ldg(reinterpret_cast<const int4*>(src) + i);
        const int w[4] = {v.x, v.y, v.z, v.w};
        float f[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) { f[2 * k] = (float)(short)(w[k] & 0xffff) * sc; f[2 * k + 1] = (float)(w[k] >> 16) * sc; }
        reinterpret_cast<float4*>(dst)[2 * i] = make_float4(f[0], f[1], f[2], f[3]);
        reinterpret_cast<float4*>(dst)[2 * i + 1] = make_float4(f[4], f[5], f[6], f[7]);
    }
    for (int i = 8 * n8 + blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) dst[i] = (float)src[i] * sc;
}

// ================================================================================================ output formats
// 8-bit quantisation <-> TGA of a row-major (n_mels, width) image (reference src/quant.rs:38-88,140-165): the step after
// the path for TGA interchange / whisper.cpp.  Byte and f32 arithmetic identical to the reference (no fused multiply-add:
// every product and sum is rounded on its own), so the bytes are bit-exact for identical f32 input.
constexpr int kTgaHeader = 26;   // 18-byte TGA header + 8-byte ID field carrying f32 min, max (src/quant.rs:44-57)

// Pass 1: per-block (min, max) partials.  grid = (nblk, n_imgs).  f32::min / f32::max ignore NaN like fminf / fmaxf.
__global__ void __launch_bounds__(256) melspec_minmax_kernel(const float* img, long long img_stride, long long n, float2* partials) {
    const float* src = img + (long long)blockIdx.y * img_stride;
    float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float v = __ldg(src + i);
        mn = fminf(mn, v); mx = fmaxf(mx, v);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    __shared__ float2 red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = make_float2(mn, mx);
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { mn = fminf(mn, red[w].x); mx = fmaxf(mx, red[w].y); }
        partials[(long long)blockIdx.y * gridDim.x + blockIdx.x] = make_float2(mn, mx);
    }
}

__device__ __forceinline__ unsigned quantize_px(float v, float mn, float scale) {
    // ((value - min) * scale).round().max(0.0).min(255.0) as u8     (src/quant.rs:146-149; round = half away from zero)
    const float q = fminf(fmaxf(roundf(__fmul_rn(__fsub_rn(v, mn), scale)), 0.0f), 255.0f);
    return (unsigned)q;   // NaN was turned into 0 by fmaxf, as f32::max does
}

// Pass 2: header + pixels.  grid = (nblk2, n_imgs); every block re-reduces the image's partials (<= 256 float2 from L2).
__global__ void __launch_bounds__(256) melspec_quantize_kernel(const float* img, long long img_stride, long long n, const float2* partials,
                                                               int nblk, unsigned char* tga, long long tga_stride, int height, int width) {
    __shared__ float2 red[8];
    __shared__ float s_mn, s_scale;
    const float* src = img + (long long)blockIdx.y * img_stride;
    unsigned char* dst = tga + (long long)blockIdx.y * tga_stride;
    float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
    for (int i = threadIdx.x; i < nblk; i += 256) {
        const float2 v = partials[(long long)blockIdx.y * nblk + i];
        mn = fminf(mn, v.x); mx = fmaxf(mx, v.y);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = make_float2(mn, mx);
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { mn = fminf(mn, red[w].x); mx = fmaxf(mx, red[w].y); }
        s_mn = mn;
        s_scale = __fdiv_rn(255.0f, __fsub_rn(mx, mn));   // src/quant.rs:144
        if (blockIdx.x == 0) {   // src/quant.rs:44-57
            const unsigned char hdr[18] = {8, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, (unsigned char)(width & 255), (unsigned char)(width >> 8),
                                           (unsigned char)(height & 255), (unsigned char)(height >> 8), 8, 0};
            for (int i = 0; i < 18; ++i) dst[i] = hdr[i];
            const unsigned a = __float_as_uint(mn), b = __float_as_uint(mx);
            for (int i = 0; i < 4; ++i) { dst[18 + i] = (unsigned char)(a >> (8 * i)); dst[22 + i] = (unsigned char)(b >> (8 * i)); }
        }
    }
    __syncthreads();
    mn = s_mn;
    const float scale = s_scale;
    unsigned char* px = dst + kTgaHeader;
    // 32-bit stores on the aligned interior, single bytes on the (<= 3 byte) edges
    const long long head = min((long long)((4 - (reinterpret_cast<uintptr_t>(px) & 3)) & 3), n);
    const long long nwords = (n - head) / 4;
    if (blockIdx.x == 0 && threadIdx.x < 8) {
        const int i = threadIdx.x;
        if (i < head) px[i] = (unsigned char)quantize_px(__ldg(src + i), mn, scale);
        const long long tail0 = head + 4 * nwords;
        if (i >= 4 && tail0 + (i - 4) < n) px[tail0 + (i - 4)] = (unsigned char)quantize_px(__ldg(src + tail0 + (i - 4)), mn, scale);
    }
    unsigned* wdst = reinterpret_cast<unsigned*>(px + head);
    const float* wsrc = src + head;
    for (long long k = (long long)blockIdx.x * 256 + threadIdx.x; k < nwords; k += (long long)gridDim.x * 256) {
        const unsigned b0 = quantize_px(__ldg(wsrc + 4 * k), mn, scale), b1 = quantize_px(__ldg(wsrc + 4 * k + 1), mn, scale);
        const unsigned b2 = quantize_px(__ldg(wsrc + 4 * k + 2), mn, scale), b3 = quantize_px(__ldg(wsrc + 4 * k + 3), mn, scale);
        wdst[k] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
    }
}

// parse_tga_8bit + dequantize (src/quant.rs:66-88,155-165): value as f32 * ((max - min) / 255.0) + min, two roundings.
// One thread = four consecutive output floats (one float4 store when the image row is 16-byte aligned); their four bytes
// come from two aligned 32-bit loads joined by a funnel shift (the pixel data start at byte 26 of the TGA).
__global__ void __launch_bounds__(256) melspec_dequantize_kernel(const unsigned char* tga, long long tga_stride, long long n, float* img,
                                                                 long long img_stride) {
    const unsigned char* src = tga + (long long)blockIdx.y * tga_stride;
    float* dst = img + (long long)blockIdx.y * img_stride;
    __shared__ float s_rng[2];
    if (threadIdx.x < 2) {   // f32 min (bytes 18..21) and max (22..25) of the ID field, little-endian, any alignment
        unsigned v = 0;
        for (int i = 0; i < 4; ++i) v |= (unsigned)src[18 + 4 * threadIdx.x + i] << (8 * i);
        s_rng[threadIdx.x] = __uint_as_float(v);
    }
    __syncthreads();
    const float mn = s_rng[0], mx = s_rng[1];
    const float scale = __fdiv_rn(__fsub_rn(mx, mn), 255.0f);
    const unsigned char* px = src + kTgaHeader;
    const bool vec = (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
    const long long nquads = vec ? n / 4 : 0;
    for (long long k = (long long)blockIdx.x * 256 + threadIdx.x; k < nquads; k += (long long)gridDim.x * 256) {
        const uintptr_t addr = reinterpret_cast<uintptr_t>(px + 4 * k);
        const unsigned* w = reinterpret_cast<const unsigned*>(addr & ~(uintptr_t)3);
        const unsigned sh = (unsigned)(addr & 3) * 8;
        // the second word is only dereferenced when the quad really straddles it
        unsigned q;
        if (sh && k == nquads - 1) {   // last quad: its second word may end past the image, read the four bytes one by one
            const unsigned char* b = px + 4 * k;
            q = (unsigned)b[0] | ((unsigned)b[1] << 8) | ((unsigned)b[2] << 16) | ((unsigned)b[3] << 24);
        } else {
            const unsigned lo = __ldg(w), hi = sh ? __ldg(w + 1) : 0u;
            q = __funnelshift_r(lo, hi, sh);
        }
        float4 v;
        v.x = __fadd_rn(__fmul_rn((float)(q & 255u), scale), mn);
        v.y = __fadd_rn(__fmul_rn((float)((q >> 8) & 255u), scale), mn);
        v.z = __fadd_rn(__fmul_rn((float)((q >> 16) & 255u), scale), mn);
        v.w = __fadd_rn(__fmul_rn((float)(q >> 24), scale), mn);
        reinterpret_cast<float4*>(dst)[k] = v;
    }
    for (long long i = 4 * nquads + (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
        dst[i] = __fadd_rn(__fmul_rn((float)px[i], scale), mn);
}

// ================================================================================================ VAD over the mel image
// vad_boundaries (reference src/vad.rs:251-338): per column x of the row-major (height, width) image, count the rows y in
// [min(min_mel, height-2), height-2) whose 3x3 Sobel gradient (src/vad.rs:472-486) satisfies gx^2 + gy^2 >= min_energy^2;
// the column is raw-active when the count reaches min_y; then a +-4 majority vote (smooth_mask, src/vad.rs:343-360).
// f64 in the reference's operation order on the (f32-valued) pixels, so the masks are bit-exact with the CPU path.
// grid = (ceil((width-2) / 248), n_imgs), block = 256: thread j owns raw column tile0 - 4 + j, threads 4..251 smooth.
constexpr int kVadTile = 248;
__global__ void __launch_bounds__(256) melspec_vad_kernel(const float* img, long long img_stride, int height, int width, double min_energy_sq,
                                                          int min_y, int min_mel, unsigned char* raw_out, unsigned char* smooth_out,
                                                          long long mask_stride) {
    __shared__ unsigned char s_raw[256];
    const float* a = img + (long long)blockIdx.y * img_stride;
    const int n = width - 2;
    const int x = blockIdx.x * kVadTile - 4 + (int)threadIdx.x;
    unsigned char act = 0;
    if (x >= 0 && x < n) {
        if (min_y == 0) {
            act = 1;
        } else {
            const int y0 = min(min_mel, height - 2);
            int count = 0;
            const float* p = a + (long long)y0 * width + x;
            double r0l = p[0], r0c = p[1], r0r = p[2];
            double r1l = p[width], r1c = p[width + 1], r1r = p[width + 2];
            for (int y = y0; y < height - 2; ++y) {
                const float* q = a + (long long)(y + 2) * width + x;
                const double r2l = q[0], r2c = q[1], r2r = q[2];
                const double gx = __dsub_rn(__dadd_rn(__dadd_rn(r0r, __dmul_rn(2.0, r1r)), r2r), __dadd_rn(__dadd_rn(r0l, __dmul_rn(2.0, r1l)), r2l));
                const double gy = __dsub_rn(__dadd_rn(__dadd_rn(r2l, __dmul_rn(2.0, r2c)), r2r), __dadd_rn(__dadd_rn(r0l, __dmul_rn(2.0, r0c)), r0r));
                if (__dadd_rn(__dmul_rn(gx, gx), __dmul_rn(gy, gy)) >= min_energy_sq) ++count;
                r0l = r1l; r0c = r1c; r0r = r1r;
                r1l = r2l; r1c = r2c; r1r = r2r;
            }
            (void)r1c;
            act = count >= min_y ? 1 : 0;
        }
    }
    s_raw[threadIdx.x] = act;
    __syncthreads();
    if (threadIdx.x >= 4 && threadIdx.x < 4 + kVadTile && x < n) {
        if (raw_out) raw_out[(long long)blockIdx.y * mask_stride + x] = act;
        const int st = max(x - 4, 0), en = min(x + 5, n);
        int cnt = 0;
        for (int k = st; k < en; ++k) cnt += s_raw[(int)threadIdx.x + (k - x)];
        smooth_out[(long long)blockIdx.y * mask_stride + x] = (cnt * 2 >= en - st) ? 1 : 0;
    }
}

// VoiceActivityDetector::add_activity (src/vad.rs:163-207) for every frame index at once: frame i >= min_x - 1 looks at
// the window of the last min_x columns, i.e. raw columns s .. s + min_x - 3 (s = i + 1 - min_x; the raw decision of a
// column triple does not depend on the window), smooths them inside the window and reports
// (active = first column intersected, leading_active_columns, active_columns).  Frames before that: (-1, -1, -1).
__global__ void __launch_bounds__(256) melspec_vad_activity_kernel(const unsigned char* raw, long long mask_stride, int height, int width,
                                                                   int min_x, int* out, long long out_stride) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= width) return;
    int* o = out + (long long)blockIdx.y * out_stride + 3ll * i;
    if (i + 1 < min_x) { o[0] = -1; o[1] = -1; o[2] = -1; return; }
    if (height < 3 || min_x < 3) { o[0] = 0; o[1] = 0; o[2] = 0; return; }   // vad_boundaries' empty EdgeInfo (src/vad.rs:264-266)
    const unsigned char* r = raw + (long long)blockIdx.y * mask_stride + (i + 1 - min_x);
    const int n = min_x - 2;
    int active = 0, leading = 0, total = 0;
    bool lead_open = true;
    for (int j = 0; j < n; ++j) {
        const int st = max(j - 4, 0), en = min(j + 5, n);
        int cnt = 0;
        for (int k = st; k < en; ++k) cnt += r[k];
        const bool sm = cnt * 2 >= en - st;
        if (j == 0) active = sm;
        if (sm) { ++total; if (lead_open) ++leading; } else lead_open = false;
    }
    o[0] = active; o[1] = leading; o[2] = total;
}

}  // namespace melspec
