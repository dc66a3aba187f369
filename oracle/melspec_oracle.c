/*
 * CPU oracle / CPU baseline for the log-mel hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, f64 restatement of the reference's CPU algorithm (wavey-ai/mel-spec @ ac3bbdd):
 *   Whisper batch path   src/stft.rs:89-169 + src/mel.rs:13-32,48-71,148-168,547-654
 *   Kaldi fbank path     src/fbank.rs:94-132,141-236,253-313
 * The reference's FFT is the un-vendored crate rustfft ^6.2.0 (Cargo.toml:18), used strictly as an
 * unnormalised forward complex DFT; here it is a mixed-radix (4/2/5/3) Stockham FFT in f64.
 *
 * Parity status: PINNED through tests/test_oracle.py — this file must agree with
 * oracle/melspec_oracle.py (<= 1e-12 on filterbanks, identical f32 outputs up to 1 ulp-level log noise)
 * which is itself bit-exact against testdata/rust_jfk_golden.npy.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 * Like the reference (src/stft.rs:104-112) it runs a FULL complex N-point FFT per real frame in f64, so the
 * timed baseline is faithful to the reference's arithmetic, not an optimised real-FFT.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double re, im; } cpx;

/* ------------------------------------------------------------------ FFT plan */
typedef struct {
    int n;
    int nstages;
    int radix[32];
    cpx *tw[32];  /* per stage: m*(r-1) twiddles w_n^(p*j), j=1..r-1 */
} fft_plan;

static void plan_init(fft_plan *pl, int n) {
    pl->n = n;
    pl->nstages = 0;
    int rem = n;
    const int cand[4] = {4, 2, 5, 3};
    while (rem > 1) {
        int r = 0;
        for (int c = 0; c < 4; ++c)
            if (rem % cand[c] == 0) { r = cand[c]; break; }
        if (!r) { /* generic prime factor */
            for (r = 7; rem % r; r += 2) {}
        }
        pl->radix[pl->nstages++] = r;
        rem /= r;
    }
    int cur = n;
    for (int s = 0; s < pl->nstages; ++s) {
        int r = pl->radix[s], m = cur / r;
        pl->tw[s] = (cpx *)malloc(sizeof(cpx) * (size_t)m * (size_t)(r - 1));
        for (int p = 0; p < m; ++p)
            for (int j = 1; j < r; ++j) {
                double a = -2.0 * M_PI * (double)p * (double)j / (double)cur;
                pl->tw[s][p * (r - 1) + (j - 1)].re = cos(a);
                pl->tw[s][p * (r - 1) + (j - 1)].im = sin(a);
            }
        cur = m;
    }
}

static void plan_free(fft_plan *pl) {
    for (int s = 0; s < pl->nstages; ++s) free(pl->tw[s]);
}

static inline cpx cmul(cpx a, cpx b) { cpx r = {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; return r; }
static inline cpx cadd(cpx a, cpx b) { cpx r = {a.re + b.re, a.im + b.im}; return r; }
static inline cpx csub(cpx a, cpx b) { cpx r = {a.re - b.re, a.im - b.im}; return r; }
static inline cpx mulmi(cpx a) { cpx r = {a.im, -a.re}; return r; } /* a * (-i) */

/* Forward DFT, Stockham autosort (decimation in frequency).  Result ends up in `x`. `y` is scratch. */
static void fft_forward(const fft_plan *pl, cpx *x, cpx *y) {
    int n = pl->n, s = 1;
    cpx *src = x, *dst = y;
    for (int st = 0; st < pl->nstages; ++st) {
        const int r = pl->radix[st], m = n / r;
        const cpx *tw = pl->tw[st];
        if (r == 4) {
            for (int p = 0; p < m; ++p) {
                const cpx w1 = tw[p * 3], w2 = tw[p * 3 + 1], w3 = tw[p * 3 + 2];
                for (int q = 0; q < s; ++q) {
                    cpx a = src[q + s * p], b = src[q + s * (p + m)], c = src[q + s * (p + 2 * m)], d = src[q + s * (p + 3 * m)];
                    cpx apc = cadd(a, c), amc = csub(a, c), bpd = cadd(b, d), jbmd = mulmi(csub(b, d));
                    dst[q + s * (4 * p)] = cadd(apc, bpd);
                    dst[q + s * (4 * p + 1)] = cmul(w1, cadd(amc, jbmd));
                    dst[q + s * (4 * p + 2)] = cmul(w2, csub(apc, bpd));
                    dst[q + s * (4 * p + 3)] = cmul(w3, csub(amc, jbmd));
                }
            }
        } else if (r == 2) {
            for (int p = 0; p < m; ++p) {
                const cpx w1 = tw[p];
                for (int q = 0; q < s; ++q) {
                    cpx a = src[q + s * p], b = src[q + s * (p + m)];
                    dst[q + s * (2 * p)] = cadd(a, b);
                    dst[q + s * (2 * p + 1)] = cmul(w1, csub(a, b));
                }
            }
        } else if (r == 5) {
            const double c1 = 0.30901699437494745, c2 = -0.80901699437494745;   /* cos(2pi/5), cos(4pi/5) */
            const double s1 = 0.95105651629515353, s2 = 0.58778525229247314;    /* sin(2pi/5), sin(4pi/5) */
            for (int p = 0; p < m; ++p) {
                const cpx *w = tw + p * 4;
                for (int q = 0; q < s; ++q) {
                    cpx x0 = src[q + s * p], x1 = src[q + s * (p + m)], x2 = src[q + s * (p + 2 * m)],
                        x3 = src[q + s * (p + 3 * m)], x4 = src[q + s * (p + 4 * m)];
                    cpx t1 = cadd(x1, x4), t2 = cadd(x2, x3), d1 = csub(x1, x4), d2 = csub(x2, x3);
                    cpx a1 = {x0.re + c1 * t1.re + c2 * t2.re, x0.im + c1 * t1.im + c2 * t2.im};
                    cpx a2 = {x0.re + c2 * t1.re + c1 * t2.re, x0.im + c2 * t1.im + c1 * t2.im};
                    cpx b1 = {s1 * d1.re + s2 * d2.re, s1 * d1.im + s2 * d2.im};
                    cpx b2 = {s2 * d1.re - s1 * d2.re, s2 * d1.im - s1 * d2.im};
                    /* y_k = a -/+ i*b  (forward transform: y1 = a1 - i b1) */
                    cpx y1 = {a1.re + b1.im, a1.im - b1.re}, y4 = {a1.re - b1.im, a1.im + b1.re};
                    cpx y2 = {a2.re + b2.im, a2.im - b2.re}, y3 = {a2.re - b2.im, a2.im + b2.re};
                    dst[q + s * (5 * p)] = cadd(x0, cadd(t1, t2));
                    dst[q + s * (5 * p + 1)] = cmul(w[0], y1);
                    dst[q + s * (5 * p + 2)] = cmul(w[1], y2);
                    dst[q + s * (5 * p + 3)] = cmul(w[2], y3);
                    dst[q + s * (5 * p + 4)] = cmul(w[3], y4);
                }
            }
        } else { /* generic radix: O(r^2) small DFT */
            for (int p = 0; p < m; ++p)
                for (int q = 0; q < s; ++q)
                    for (int j = 0; j < r; ++j) {
                        cpx acc = {0.0, 0.0};
                        for (int k = 0; k < r; ++k) {
                            double a = -2.0 * M_PI * (double)((j * k) % r) / (double)r;
                            cpx w = {cos(a), sin(a)};
                            acc = cadd(acc, cmul(src[q + s * (p + m * k)], w));
                        }
                        dst[q + s * (r * p + j)] = j ? cmul(tw[p * (r - 1) + j - 1], acc) : acc;
                    }
        }
        cpx *t = src; src = dst; dst = t;
        n = m;
        s *= r;
    }
    if (src != x) memcpy(x, src, sizeof(cpx) * (size_t)pl->n);
}

/* ------------------------------------------------------------------ filterbanks */
static double sl_hz_to_mel(double f) { /* src/mel.rs:591-607, htk=false */
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
static double sl_mel_to_hz(double m) { /* src/mel.rs:609-625 */
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}

/* src/mel.rs:547-589 `mel(sr, n_fft, n_mels, None, None, false, true)`; out: n_mels x (n_fft/2+1) row-major */
void oracle_slaney_filterbank(double sr, int n_fft, int n_mels, double *out) {
    const int nb = n_fft / 2 + 1, np = n_mels + 2;
    double *mel_f = (double *)malloc(sizeof(double) * (size_t)np);
    const double lo = sl_hz_to_mel(0.0), hi = sl_hz_to_mel(sr / 2.0), step = (hi - lo) / (double)(np - 1);
    for (int i = 0; i < np; ++i) mel_f[i] = sl_mel_to_hz(lo + step * (double)i);
    for (int i = 0; i < n_mels; ++i) {
        const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
        const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
        for (int b = 0; b < nb; ++b) {
            const double f = (sr / (double)n_fft) * (double)b;
            double lower = -(mel_f[i] - f) / fd0, upper = (mel_f[i + 2] - f) / fd1;
            lower = fmin(fmax(lower, 0.0), 1.0);
            upper = fmin(fmax(upper, 0.0), 1.0);
            out[(size_t)i * nb + b] = fmin(lower, upper) * enorm;
        }
    }
    free(mel_f);
}

/* src/fbank.rs:253-313 */
void oracle_kaldi_filterbank(double sr, int fft_size, int n_mels, double low, double high, double *out) {
    const int nb = fft_size / 2 + 1;
    const double ml = 1127.0 * log(1.0 + low / 700.0), mh = 1127.0 * log(1.0 + high / 700.0);
    double *hz = (double *)malloc(sizeof(double) * (size_t)(n_mels + 2));
    for (int i = 0; i < n_mels + 2; ++i)
        hz[i] = 700.0 * (exp((ml + (mh - ml) * (double)i / (double)(n_mels + 1)) / 1127.0) - 1.0);
    memset(out, 0, sizeof(double) * (size_t)n_mels * nb);
    for (int m = 0; m < n_mels; ++m) {
        const double l = hz[m], c = hz[m + 1], r = hz[m + 2];
        if (c <= l || r <= c) continue;
        for (int b = 0; b < nb; ++b) {
            const double f = (double)b * sr / (double)fft_size;
            if (f > l && f <= c) out[(size_t)m * nb + b] = (f - l) / (c - l);
            else if (f > c && f < r) out[(size_t)m * nb + b] = (r - f) / (r - c);
        }
    }
    free(hz);
}

typedef struct { int n_mels, nb; int *start, *len; double *w; /* w: rows packed, row m at off[m] */ int *off; } sparse_bank;

static void sparse_from_dense(sparse_bank *sb, const double *dense, int n_mels, int nb) { /* src/mel.rs:48-71 */
    sb->n_mels = n_mels; sb->nb = nb;
    sb->start = (int *)malloc(sizeof(int) * (size_t)n_mels * 3);
    sb->len = sb->start + n_mels; sb->off = sb->len + n_mels;
    sb->w = (double *)malloc(sizeof(double) * (size_t)n_mels * nb * 2);
    /* rows are stored as (bin, weight) pairs to allow gaps, like the reference */
    int o = 0;
    for (int m = 0; m < n_mels; ++m) {
        sb->off[m] = o; int cnt = 0;
        for (int b = 0; b < nb; ++b)
            if (dense[(size_t)m * nb + b] != 0.0) { sb->w[2 * (o + cnt)] = (double)b; sb->w[2 * (o + cnt) + 1] = dense[(size_t)m * nb + b]; ++cnt; }
        sb->len[m] = cnt; sb->start[m] = 0; o += cnt;
    }
}
static void sparse_free(sparse_bank *sb) { free(sb->start); free(sb->w); }

/* ------------------------------------------------------------------ per-clip kernels */
/* Whisper: src/stft.rs:119-138.  out: F x n_mels f32 (frame-major). Returns frames. */
static int64_t whisper_clip(const float *x, int64_t n, int fft, int hop, const sparse_bank *sb, const double *win,
                            const fft_plan *pl, cpx *buf, cpx *scr, double *mel, float *out) {
    if (n < fft) return 0;
    const int64_t nf = (n - fft) / hop + 1;
    const int half = fft / 2, M = sb->n_mels;
    for (int64_t k = 0; k < nf; ++k) {
        const float *fr = x + k * hop;
        for (int i = 0; i < fft; ++i) { buf[i].re = (double)fr[i] * win[i]; buf[i].im = 0.0; } /* stft.rs:147-169,104-108 */
        fft_forward(pl, buf, scr);                                                              /* stft.rs:110 */
        double mx = -INFINITY;
        for (int m = 0; m < M; ++m) {                                                           /* mel.rs:148-168 */
            double e = 0.0;
            const double *row = sb->w + 2 * sb->off[m];
            for (int j = 0; j < sb->len[m]; ++j) {
                const int b = (int)row[2 * j];
                const double p = b < half ? buf[b].re * buf[b].re + buf[b].im * buf[b].im : 0.0;
                e += row[2 * j + 1] * p;
            }
            mel[m] = log10(e > 1e-10 ? e : 1e-10);
            if (mel[m] > mx) mx = mel[m];
        }
        mx -= 8.0;                                                                              /* mel.rs:645-654 */
        for (int m = 0; m < M; ++m) out[k * M + m] = (float)(((mel[m] > mx ? mel[m] : mx) + 4.0) / 4.0);
    }
    return nf;
}

/* Kaldi: src/fbank.rs:141-236 with the default config (power, log, preemph 0.97, CMN). */
static int64_t kaldi_clip(const float *x, int64_t n, int frame_len, int shift, int fft, double preemph, int cmn,
                          const sparse_bank *sb, const double *win, const fft_plan *pl, cpx *buf, cpx *scr,
                          double *fbuf, double *pw, float *out) {
    if (n < frame_len) return 0;
    const int64_t T = 1 + (n - frame_len) / shift;
    const int M = sb->n_mels, nb = fft / 2 + 1;
    for (int64_t k = 0; k < T; ++k) {
        const int64_t start = k * shift;
        double mean = 0.0;
        for (int i = 0; i < frame_len; ++i) mean += (double)x[start + i];
        mean /= (double)frame_len;
        for (int i = 0; i < frame_len; ++i) fbuf[i] = (double)x[start + i] - mean;
        if (preemph > 0.0) {
            for (int i = frame_len - 1; i >= 1; --i) fbuf[i] -= preemph * fbuf[i - 1];
            if (start > 0) fbuf[0] -= preemph * ((double)x[start - 1] - mean);
        }
        for (int i = 0; i < frame_len; ++i) { buf[i].re = fbuf[i] * win[i]; buf[i].im = 0.0; }
        for (int i = frame_len; i < fft; ++i) { buf[i].re = 0.0; buf[i].im = 0.0; }
        fft_forward(pl, buf, scr);
        for (int b = 0; b < nb; ++b) pw[b] = buf[b].re * buf[b].re + buf[b].im * buf[b].im;
        for (int m = 0; m < M; ++m) {
            double e = 0.0;
            const double *row = sb->w + 2 * sb->off[m];
            for (int j = 0; j < sb->len[m]; ++j) e += row[2 * j + 1] * pw[(int)row[2 * j]];
            const double fl = 1.1920928955078125e-07; /* f32::EPSILON */
            e = e > fl ? e : fl;
            out[k * M + m] = (float)log(e);
        }
    }
    if (cmn && T > 0) { /* fbank.rs:226-233: f32 mean per mel bin over time */
        for (int m = 0; m < M; ++m) {
            /* ndarray's f32 mean: pairwise-style sum; we use a f64 accumulator rounded to f32, which is within
               1 ulp of any f32 summation order's exact value — the test tolerance covers the difference. */
            double acc = 0.0;
            for (int64_t k = 0; k < T; ++k) acc += (double)out[k * M + m];
            const float mean = (float)(acc / (double)T);
            for (int64_t k = 0; k < T; ++k) out[k * M + m] -= mean;
        }
    }
    return T;
}

/* ------------------------------------------------------------------ batch drivers (clips sharded over threads) */
typedef struct {
    int kind; /* 0 whisper, 1 kaldi */
    const float *pcm; int64_t n_clips, stride, n_samples;
    int fft, hop, n_mels, frame_len; double sr, preemph; int cmn;
    float *out; int64_t frames_per_clip;
    int tid, nthreads;
    const sparse_bank *sb; const double *win;
} job;

static void *worker(void *arg) {
    job *j = (job *)arg;
    fft_plan pl; plan_init(&pl, j->fft);
    cpx *buf = (cpx *)malloc(sizeof(cpx) * (size_t)j->fft * 2), *scr = buf + j->fft;
    double *tmp = (double *)malloc(sizeof(double) * (size_t)(j->n_mels + j->fft * 2 + 8));
    for (int64_t c = j->tid; c < j->n_clips; c += j->nthreads) {
        const float *x = j->pcm + c * j->stride;
        float *o = j->out + c * j->frames_per_clip * j->n_mels;
        if (j->kind == 0) whisper_clip(x, j->n_samples, j->fft, j->hop, j->sb, j->win, &pl, buf, scr, tmp, o);
        else kaldi_clip(x, j->n_samples, j->frame_len, j->hop, j->fft, j->preemph, j->cmn, j->sb, j->win, &pl, buf, scr,
                        tmp, tmp + j->fft, o);
    }
    free(tmp); free(buf); plan_free(&pl);
    return NULL;
}

static void run_jobs(job *proto, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256]; job jobs[256];
    for (int t = 0; t < threads; ++t) { jobs[t] = *proto; jobs[t].tid = t; jobs[t].nthreads = threads; }
    if (threads == 1) { worker(&jobs[0]); return; }
    for (int t = 0; t < threads; ++t) pthread_create(&th[t], NULL, worker, &jobs[t]);
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
}

int64_t oracle_num_frames(int64_t n, int fft, int hop) { return n < fft ? 0 : (n - fft) / hop + 1; }

/* out: [n_clips][F][n_mels] f32.  Returns frames per clip. */
int64_t oracle_whisper_batch(const float *pcm, int64_t n_clips, int64_t stride, int64_t n_samples, int fft, int hop,
                             int n_mels, double sr, float *out, int threads) {
    const int nb = fft / 2 + 1;
    double *dense = (double *)malloc(sizeof(double) * (size_t)n_mels * nb);
    oracle_slaney_filterbank(sr, fft, n_mels, dense);
    sparse_bank sb; sparse_from_dense(&sb, dense, n_mels, nb);
    double *win = (double *)malloc(sizeof(double) * (size_t)fft);
    for (int i = 0; i < fft; ++i) win[i] = 0.5 * (1.0 - cos((2.0 * M_PI * (double)i) / (double)fft)); /* stft.rs:141-145 */
    job j; memset(&j, 0, sizeof j);
    j.kind = 0; j.pcm = pcm; j.n_clips = n_clips; j.stride = stride; j.n_samples = n_samples; j.fft = fft; j.hop = hop;
    j.n_mels = n_mels; j.sr = sr; j.out = out; j.frames_per_clip = oracle_num_frames(n_samples, fft, hop); j.sb = &sb; j.win = win;
    run_jobs(&j, threads);
    free(win); sparse_free(&sb); free(dense);
    return j.frames_per_clip;
}

/* Kaldi default config (src/fbank.rs:46-64) at 16 kHz-like rates.  out: [n_clips][T][n_mels] f32. */
int64_t oracle_kaldi_batch(const float *pcm, int64_t n_clips, int64_t stride, int64_t n_samples, double sr, int n_mels,
                           double frame_ms, double shift_ms, double preemph, double low, double high, int cmn,
                           float *out, int threads) {
    const int frame_len = (int)round(frame_ms / 1000.0 * sr), shift = (int)round(shift_ms / 1000.0 * sr);
    int fft = 1; while (fft < frame_len) fft <<= 1;
    const int nb = fft / 2 + 1;
    double *dense = (double *)malloc(sizeof(double) * (size_t)n_mels * nb);
    oracle_kaldi_filterbank(sr, fft, n_mels, low, high == 0.0 ? sr / 2.0 : high, dense);
    sparse_bank sb; sparse_from_dense(&sb, dense, n_mels, nb);
    double *win = (double *)malloc(sizeof(double) * (size_t)frame_len);
    for (int i = 0; i < frame_len; ++i) win[i] = pow(0.5 - 0.5 * cos(2.0 * M_PI * (double)i / (double)(frame_len - 1)), 0.85);
    job j; memset(&j, 0, sizeof j);
    j.kind = 1; j.pcm = pcm; j.n_clips = n_clips; j.stride = stride; j.n_samples = n_samples; j.fft = fft; j.hop = shift;
    j.frame_len = frame_len; j.n_mels = n_mels; j.sr = sr; j.preemph = preemph; j.cmn = cmn; j.out = out;
    j.frames_per_clip = n_samples < frame_len ? 0 : 1 + (n_samples - frame_len) / shift; j.sb = &sb; j.win = win;
    run_jobs(&j, threads);
    free(win); sparse_free(&sb); free(dense);
    return j.frames_per_clip;
}

/* Raw FFT, exported so the tests can check the mixed-radix transform against numpy. in/out interleaved re,im. */
void oracle_fft_forward(const double *in, double *out, int n) {
    fft_plan pl; plan_init(&pl, n);
    cpx *buf = (cpx *)malloc(sizeof(cpx) * (size_t)n * 2);
    memcpy(buf, in, sizeof(cpx) * (size_t)n);
    fft_forward(&pl, buf, buf + n);
    memcpy(out, buf, sizeof(cpx) * (size_t)n);
    free(buf); plan_free(&pl);
}
