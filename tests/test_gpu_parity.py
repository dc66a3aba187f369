"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Needs a B200: `pytest -m gpu`.

Tolerance (BASELINE.json north_star / SURVEY §8c): Whisper path <= 1e-4 max-abs against the f64 oracle, written
as WHISPER_TOL below.  The reference's own CUDA-vs-CPU test accepts 0.08 max / 0.01 mean (src/cuda.rs:540-544).
"""
import os

import numpy as np
import pytest

import melspec_oracle as o
import oracle_c as oc

pytestmark = pytest.mark.gpu

WHISPER_TOL = 1e-4


@pytest.fixture(scope="module")
def m():
    import mel_spec_b200 as mod
    mod.build()
    return mod


@pytest.fixture(scope="module")
def torch():
    import torch as t
    assert t.cuda.is_available(), "these tests need a GPU"
    return t


@pytest.fixture(scope="module")
def mel400(m):
    h = m.CudaMelSpectrogram(400, 160, 16000.0, 80)
    yield h
    h.close()


def _device_run(torch, h, pcm, lens=None, layout=0, n_mels=80):
    x = torch.from_numpy(np.ascontiguousarray(pcm)).cuda()
    b, s = x.shape
    f = h.num_frames(s)
    shape = (b, f, n_mels) if layout == 0 else (b, n_mels, f)
    out = torch.full(shape, float("nan"), dtype=torch.float32, device="cuda")
    dl = None if lens is None else torch.tensor(lens, dtype=torch.int32, device="cuda")
    h.compute_device(x, b, s, s, out, d_lens=dl, layout=layout)
    torch.cuda.synchronize()
    return out.cpu().numpy()


# ------------------------------------------------------------------------------------------ reference's own cases
def test_jfk_whisper400_vs_oracle(m, mel400, jfk):
    # BASELINE config 1 at the Whisper configuration (fft 400): no golden file exists for fft 400,
    # the pinned oracle is the truth.
    got = mel400.compute_mel_spectrogram(jfk)
    want = o.whisper_mel_batch(jfk, 400, 160, 80, 16000.0)
    assert got.shape == want.shape == (1098, 80)
    d = np.abs(got - want)
    assert d.max() <= WHISPER_TOL, d.max()


def test_reference_cuda_test_signal(m, mel400):
    # src/cuda.rs:488-545: 1 s of 4 sines; same frame count; reference tolerance 0.08/0.01, ours 1e-4
    x = o.reference_test_signal()
    got = mel400.compute_mel_spectrogram(x)
    want = o.whisper_mel_batch(x)
    assert got.shape == want.shape == (98, 80)
    d = np.abs(got - want)
    assert d.max() < 0.08 and d.mean() < 0.01
    assert d.max() <= WHISPER_TOL, d.max()


def test_readme_shapes_and_silence(m, mel400):
    # tests/readme_examples.rs:11-18: zeros(16000) -> non-empty, rows of 80; silence sits on the 1e-10 floor => -1.5
    got = mel400.compute_mel_spectrogram(np.zeros(16000, np.float32))
    assert got.shape == (98, 80)
    assert np.all(got == -1.5)
    got2 = m.Spectrogram.compute_mel_spectrogram(np.zeros(16000, np.float32), 400, 160, 80, 16000.0)
    assert np.array_equal(got, got2)


def test_empty_and_short_inputs(m, mel400):
    # src/cuda.rs:91-93: too-short input => Ok(vec![])
    for n in (0, 1, 399):
        assert mel400.compute_mel_spectrogram(np.zeros(n, np.float32)).shape == (0, 80)
    got = mel400.compute_mel_spectrogram(np.ones(400, np.float32) * 0.5)
    want = o.whisper_mel_batch(np.ones(400, np.float32) * 0.5)
    assert got.shape == (1, 80) and np.abs(got - want).max() <= WHISPER_TOL
    assert mel400.max_frames_per_batch() == 8192          # src/cuda.rs:150-155 at fft 400 / 80 mels


# ------------------------------------------------------------------------------------------ synthetic batches
def test_synthetic_batch_vs_oracle(m, mel400, torch):
    # BASELINE config 2 shape at a size the oracle finishes in seconds: first 8 clips of the 10 s workload
    pcm = np.stack([o.synth_clip(i, 160000) for i in range(8)])
    got = _device_run(torch, mel400, pcm)
    want = oc.whisper_batch(pcm, threads=8)
    assert got.shape == want.shape == (8, 998, 80)
    d = np.abs(got - want)
    assert d.max() <= WHISPER_TOL, d.max()
    # clip 7 starts with one second of digital silence: floor frames are exactly -1.5
    assert np.all(got[7, :90] == -1.5)


@pytest.mark.parametrize("n", [400, 559, 560, 1199, 7919, 8080, 16000, 16001, 16002, 16003, 30000])
def test_ragged_lengths_single_clip(m, mel400, n):
    # every tail shape: partial last tile, partial last warp pass, odd sample counts (non-TMA path on the host side
    # pads rows to 4 samples; the device path is exercised unaligned in test_unaligned_device_pointers)
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) * 0.1).astype(np.float32)
    got = mel400.compute_mel_spectrogram(x)
    want = o.whisper_mel_batch(x)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= WHISPER_TOL


def test_per_clip_lengths(m, mel400, torch):
    rng = np.random.default_rng(5)
    s = 20000
    lens = [20000, 0, 399, 400, 4000, 12345, 19999, 8080]
    pcm = (rng.standard_normal((len(lens), s)) * 0.2).astype(np.float32)
    got = _device_run(torch, mel400, pcm, lens=lens)
    for i, n in enumerate(lens):
        want = o.whisper_mel_batch(pcm[i, :n])
        f = want.shape[0]
        assert np.abs(got[i, :f] - want).max() <= WHISPER_TOL if f else True
        assert np.isnan(got[i, f:]).all(), "frames past a clip's own length must stay untouched"


def test_mel_major_layout(m, mel400, torch):
    pcm = np.stack([o.synth_clip(i, 16000) for i in range(3)])
    a = _device_run(torch, mel400, pcm, layout=0)
    b = _device_run(torch, mel400, pcm, layout=1)
    assert b.shape == (3, 80, 98)
    assert np.array_equal(a.transpose(0, 2, 1), b)


@pytest.mark.parametrize("n_mels", [80, 128])
def test_mel_major_large_batch_equals_frame_major(m, torch, n_mels):
    """The mel-major launch shapes of round 2 (CTA-contiguous tile ranges with interleaved warps, offset-table stores, compiled-in
    80 / 128-mel schedules; a ragged last tile and a batch large enough that every warp owns several tiles): bit-identical to the
    frame-major output transposed, which is checked against the oracle; untouched padding behind a wider row stride stays untouched."""
    h = m.CudaMelSpectrogram(400, 160, 16000.0, n_mels)
    clips, n = 300, 16000 * 2 + 400 + 160 * 3                     # 204 frames = 34 tiles of 6; 300 clips > 148 CTAs x 12 warps / 34
    pcm = np.stack([o.synth_clip(i % 7, n) * (0.1 + 0.05 * (i % 5)) for i in range(clips)]).astype(np.float32)
    f = h.num_frames(n)
    assert f == 204
    a = _device_run(torch, h, pcm, layout=0, n_mels=n_mels)
    b = _device_run(torch, h, pcm, layout=1, n_mels=n_mels)
    assert np.array_equal(a.transpose(0, 2, 1), b)
    for i in (0, 1, 299):
        assert np.abs(a[i] - o.whisper_mel_batch(pcm[i], 400, 160, n_mels, 16000.0)).max() <= WHISPER_TOL
    pcm2 = np.ascontiguousarray(pcm[:, : n - 160 * 2])              # 202 frames: ragged last tile (4 of 6 frames)
    a2 = _device_run(torch, h, pcm2, layout=0, n_mels=n_mels)
    b2 = _device_run(torch, h, pcm2, layout=1, n_mels=n_mels)
    assert a2.shape[1] == 202 and np.array_equal(a2.transpose(0, 2, 1), b2)
    # interleave_frames image with padding columns (even width > frames): columns [frames, width) are zeros
    x = torch.from_numpy(pcm2).cuda()
    img = torch.full((clips, n_mels, 210), float("nan"), dtype=torch.float32, device="cuda")
    h.compute_interleaved_device(x, clips, pcm2.shape[1], pcm2.shape[1], 210, img)
    torch.cuda.synchronize()
    got = img.cpu().numpy()
    assert np.array_equal(got[:, :, :202], b2) and np.all(got[:, :, 202:] == 0.0)
    h.close()


@pytest.mark.parametrize("n_mels,layout", [(128, 1), (80, 1), (80, 0)])
def test_repeated_launches_are_identical(m, torch, n_mels, layout):
    """Regression test of the single-buffered PCM stage (profiles/r2_refill_guard.md): before the proxy fence between a pass's sample
    loads and the TMA refill of the same buffer, about 1 launch in 100 of this ragged mel-major batch computed one frame from the
    next tile's samples.  300 launches of the same input, each compared with the first on the device, bit for bit."""
    h = m.CudaMelSpectrogram(400, 160, 16000.0, n_mels)
    clips, n = 300, 400 + 160 * 201                                   # 202 frames: 33 full tiles + a ragged one
    pcm = np.stack([o.synth_clip(i % 7, n) * (0.1 + 0.05 * (i % 5)) for i in range(clips)]).astype(np.float32)
    x = torch.from_numpy(pcm).cuda()
    shape = (clips, 202, n_mels) if layout == 0 else (clips, n_mels, 202)
    ref = torch.empty(shape, dtype=torch.float32, device="cuda")
    h.compute_device(x, clips, n, n, ref, layout=layout)
    torch.cuda.synchronize()
    differing = 0
    for _ in range(300):
        out = torch.full(shape, float("nan"), dtype=torch.float32, device="cuda")
        h.compute_device(x, clips, n, n, out, layout=layout)
        differing += int(bool((out != ref).any()))
    assert differing == 0
    h.close()


def test_unaligned_device_pointers(m, mel400, torch):
    # odd strides / offsets force the cooperative-copy input path and the plain-store output path
    rng = np.random.default_rng(9)
    s = 8003
    pcm = (rng.standard_normal((3, s)) * 0.3).astype(np.float32)
    buf = torch.zeros(3 * s + 1, dtype=torch.float32, device="cuda")
    buf[1:] = torch.from_numpy(pcm.reshape(-1)).cuda()
    f = mel400.num_frames(s)
    outbuf = torch.full((3 * f * 80 + 1,), float("nan"), dtype=torch.float32, device="cuda")
    mel400.compute_device(buf[1:], 3, s, s, outbuf[1:])
    torch.cuda.synchronize()
    got = outbuf[1:].cpu().numpy().reshape(3, f, 80)
    want = np.stack([o.whisper_mel_batch(pcm[i]) for i in range(3)])
    assert np.abs(got - want).max() <= WHISPER_TOL


def test_other_mel_counts_and_hops(m, torch):
    rng = np.random.default_rng(3)
    x = (rng.standard_normal(24000) * 0.2).astype(np.float32)
    for n_mels, hop in ((128, 160), (40, 160), (80, 200), (80, 128)):
        h = m.CudaMelSpectrogram(400, hop, 16000.0, n_mels)
        got = h.compute_mel_spectrogram(x)
        want = o.whisper_mel_batch(x, 400, hop, n_mels, 16000.0)
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= WHISPER_TOL, (n_mels, hop)
        h.close()


# ------------------------------------------------------------------------------------------ full-size properties
def test_full_size_batch_properties(m, mel400, torch):
    """BASELINE config 2 at full size (1024 x 10 s): size-independent properties instead of an oracle pass.
    (1) a clip's features do not depend on the batch around it; (2) every clip is finite and within the normalised
    range; (3) shifting a clip by k hops shifts its frames by k (frames are position independent); (4) spot-check
    of clips against the oracle."""
    base = np.stack([o.synth_clip(i, 160000) for i in range(16)])
    pcm = np.tile(base, (64, 1))                     # 1024 clips
    x = torch.from_numpy(pcm).cuda()
    out = torch.empty((1024, 998, 80), dtype=torch.float32, device="cuda")
    mel400.compute_device(x, 1024, 160000, 160000, out)
    torch.cuda.synchronize()
    ref = out[:16]
    for r in range(1, 64):
        assert torch.equal(out[16 * r:16 * (r + 1)], ref), r          # (1) bit-identical regardless of batch position
    got = ref.cpu().numpy()
    assert np.isfinite(got).all() and got.min() >= -1.5 and got.max() < 3.0          # (2)
    shifted = _device_run(torch, mel400, base[:4, 160 * 7:])
    assert np.abs(shifted - got[:4, 7:7 + shifted.shape[1]]).max() <= 5e-5                 # (3) Re-slot vs Im-slot rounding
    want = oc.whisper_batch(base[[0, 7, 15]], threads=3)
    assert np.abs(got[[0, 7, 15]] - want).max() <= WHISPER_TOL                              # (4)


def test_host_batch_api_matches_device_api(m, mel400, torch):
    pcm = np.stack([o.synth_clip(i, 48000) for i in range(7)])
    a = mel400.compute_host(pcm)
    b = _device_run(torch, mel400, pcm)
    assert np.array_equal(a, b)


# ------------------------------------------------------------------------------------------ streaming
def test_streaming_matches_stream_oracle(m, jfk):
    # RingBuffer semantics (src/rb.rs:86-121) at fft 400: first frame at sample 80, trailing partial hop dropped
    rb = m.RingBuffer(m.MelConfig(400, 160, 80, 16000.0), capacity=1 << 20)
    frames = []
    rng = np.random.default_rng(0)
    pos = 0
    x = jfk[:48000 + 77]
    while pos < x.size:
        n = int(rng.integers(1, 1500))
        rb.add_frame(x[pos:pos + n])
        pos += n
        while True:
            fr = rb.maybe_mel()
            if fr is None:
                break
            assert fr.shape == (80, 1)
            frames.append(fr[:, 0])
    got = np.stack(frames)
    want = o.whisper_mel_stream(x, 400, 160, 80, 16000.0)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= WHISPER_TOL
    rb.close()


def test_streaming_large_chunks(m, jfk):
    rb = m.RingBuffer(m.MelConfig(400, 160, 80, 16000.0), capacity=1 << 22, max_chunk_samples=16000)
    rb.add_frame(jfk)
    got = rb.drain()
    want = o.whisper_mel_stream(jfk, 400, 160, 80, 16000.0)
    assert got.shape == want.shape == (1098, 80)
    assert np.abs(got - want).max() <= WHISPER_TOL
    rb.close()


# ------------------------------------------------------------------------------------------ fft 512 (golden-file config)
def test_whisper512_matches_rust_golden(m, jfk, golden_dir, torch):
    """The reference's one numeric golden (src/rb.rs:134-179): fft 512 / hop 160 / 80 mels, stream framing.
    Equivalent batch form: samples[128:], mel-major (80, 1097).  Reference tolerance 1e-6 (f64 path); ours 1e-4 (fp32)."""
    gold = np.load(os.path.join(golden_dir, "rust_jfk_golden.npy"))
    h = m.CudaMelSpectrogram(512, 160, 16000.0, 80)
    got = h.compute_host(jfk[128:], layout=m.LAYOUT_MEL_MAJOR)
    assert got.shape == gold.shape == (80, 1097)
    assert np.abs(got - gold).max() <= WHISPER_TOL, np.abs(got - gold).max()
    fm = h.compute_mel_spectrogram(jfk[128:])
    assert np.array_equal(fm.T, got)
    want = o.whisper_mel_batch(jfk, 512, 160, 80, 16000.0)
    assert np.abs(h.compute_mel_spectrogram(jfk) - want).max() <= WHISPER_TOL
    h.close()


def test_ringbuffer_512_matches_rust_golden(m, jfk, golden_dir):
    # the reference test itself: RingBuffer(MelConfig(512,160,80,16k)) fed the whole file, frames interleaved mel-major
    gold = np.load(os.path.join(golden_dir, "rust_jfk_golden.npy"))
    rb = m.RingBuffer(m.MelConfig(512, 160, 80, 16000.0), capacity=1 << 22, max_chunk_samples=32000)
    rb.add_frame(jfk)
    got = rb.drain()
    assert got.T.shape == gold.shape
    assert np.abs(got.T - gold).max() <= WHISPER_TOL
    rb.close()


def test_whisper512_ragged_and_batch(m, torch):
    h = m.CudaMelSpectrogram(512, 160, 16000.0, 80)
    for n in (511, 512, 671, 672, 5000, 16003):
        rng = np.random.default_rng(n)
        x = (rng.standard_normal(n) * 0.1).astype(np.float32)
        got = h.compute_mel_spectrogram(x)
        want = o.whisper_mel_batch(x, 512, 160, 80, 16000.0)
        assert got.shape == want.shape
        if want.size:
            assert np.abs(got - want).max() <= WHISPER_TOL, n
    pcm = np.stack([o.synth_clip(i, 48000) for i in range(5)])
    got = _device_run(torch, h, pcm)
    want = oc.whisper_batch(pcm, 512, 160, 80, 16000.0, threads=4)
    assert np.abs(got - want).max() <= WHISPER_TOL
    h.close()


# ------------------------------------------------------------------------------------------ Kaldi fbank (src/fbank.rs)
# fp32 tolerance for the Kaldi path, stated: ln() is not clamped relative to the frame maximum, so low-energy bins carry
# the fp32 FFT noise floor (an all-fp32 pipeline is 2.7e-3 max on JFK).  Bar: max-abs <= 5e-3 and >= 99.5 %
# of the values within 1e-3 of the f64 oracle of the reference's semantics.
KALDI_TOL_MAX = 5e-3
KALDI_TOL_BULK = 1e-3


def _kaldi_check(got, want):
    assert got.shape == want.shape
    d = np.abs(got - want)
    assert d.max() <= KALDI_TOL_MAX, d.max()
    assert (d <= KALDI_TOL_BULK).mean() >= 0.995, (d <= KALDI_TOL_BULK).mean()


def test_kaldi_fbank_jfk(m, jfk, golden_dir):
    fb = m.Fbank(m.FbankConfig())
    got = fb.compute(jfk)
    assert got.shape == (1098, 80)                                   # src/fbank.rs:484-490 (the reference's assertion)
    assert np.isfinite(got).all() and got.var() > 0.1                 # src/fbank.rs:521-534
    _kaldi_check(got, o.kaldi_fbank(jfk))
    gold = np.load(os.path.join(golden_dir, "kaldi_fbank_jfk.npy")).T  # kaldi_native_fbank output: the reference itself
    d = np.abs(got - gold)                                             # is 1.5e-2 away from it
    assert d.max() < 2.5e-2 and d.mean() < 4e-3
    fb.close()


def test_kaldi_without_cmn_and_edge_cases(m):
    cfg = m.FbankConfig(apply_cmn=False)
    fb = m.Fbank(cfg)
    rng = np.random.default_rng(1)
    for n in (399, 400, 559, 560, 4000, 16001):
        x = (rng.standard_normal(n) * 0.05).astype(np.float32)
        got = fb.compute(x)
        want = o.kaldi_fbank(x, apply_cmn=False)
        assert got.shape == want.shape
        if want.size:
            _kaldi_check(got, want)
    assert fb.compute(np.zeros(16000, np.float32)).shape == (98, 80)   # tests/readme_examples.rs:21-31
    fb.close()


def test_kaldi_batch_vs_oracle(m, torch):
    # BASELINE config 3 shape at oracle-friendly size
    pcm = np.stack([o.synth_clip(i, 160000) for i in range(4)])
    fb = m.Fbank(m.FbankConfig())
    got = _device_run(torch, fb, pcm)
    want = oc.kaldi_batch(pcm, threads=4)
    for i in range(4):
        _kaldi_check(got[i], want[i])
    fb.close()


def test_kaldi_fused_cmn_large_batch(m, torch):
    """Batches with >= one clip per SM and >= 24 warp tiles per clip take the fused-CMN path of the plan-512 kernel (one CTA
    per clip at a time, mean subtracted in place while the clip is in L2).  Checked against the oracle on a few clips, against
    the un-normalised features minus their column means for every clip, for ragged per-clip lengths, and for determinism."""
    n_clips, n = 300, 24000                                       # 148 frames = 37 tiles of 4 per clip
    base = np.stack([o.synth_clip(i, n) for i in range(6)]).astype(np.float32)
    rng = np.random.default_rng(3)
    pcm = np.ascontiguousarray(base[rng.integers(0, 6, n_clips)] * rng.uniform(0.2, 1.0, (n_clips, 1)).astype(np.float32))
    lens = rng.integers(300, n + 1, n_clips).astype(np.int32)
    lens[:4] = n
    lens[4] = 399                                                 # a clip with no frame at all
    fb = m.Fbank(m.FbankConfig())
    raw = m.Fbank(m.FbankConfig(apply_cmn=False))
    got = _device_run(torch, fb, pcm, lens=lens.tolist())
    again = _device_run(torch, fb, pcm, lens=lens.tolist())
    plain = _device_run(torch, raw, pcm, lens=lens.tolist())
    for i in range(n_clips):
        f = 0 if lens[i] < 400 else (int(lens[i]) - 400) // 160 + 1
        if f == 0:
            assert np.isnan(got[i]).all()                         # untouched (the test buffer is NaN-filled)
            continue
        want = plain[i, :f] - plain[i, :f].mean(axis=0, dtype=np.float32)
        assert np.abs(got[i, :f] - want).max() <= 2e-5, i
        assert np.isnan(got[i, f:]).all()
        assert np.array_equal(got[i, :f], again[i, :f])
    for i in range(4):
        _kaldi_check(got[i], o.kaldi_fbank(pcm[i]))
    fb.close(); raw.close()


# ------------------------------------------------------------------------------------------ long stream (BASELINE config 5 shape)
def test_long_stream_chunked_equals_batch(m, torch):
    """10 minutes of audio through the streaming C ABI in 1 s pushes and in one large push: frame count follows the
    RingBuffer rule (whole hops only, first frame at sample 80), both chunkings agree bit for bit (the pipeline cuts
    pushes at the same 4 s pieces only if the tails line up, so equality is checked against the batch path with the
    A/B-slot tolerance), and a slice is checked against the oracle."""
    import ctypes as C
    n = 16000 * 600 + 123
    x = np.concatenate([o.synth_clip(i, 160000) for i in range(60)] + [np.zeros(123, np.float32)])[:n]
    h = m.CudaMelSpectrogram(400, 160, 16000.0, 80)
    L = m.lib()
    want_frames = n // 160 - 3 + 1
    outs = []
    for chunk in (16000, n):
        s = C.c_void_p()
        assert L.melspec_stream_create(h._h, chunk, C.byref(s)) == 0
        out = np.empty((want_frames + 8, 80), np.float32)
        got = 0
        for off in range(0, n, chunk):
            piece = np.ascontiguousarray(x[off:off + chunk])
            em = C.c_int64(0)
            rc = L.melspec_stream_push(s, piece.ctypes.data, piece.size, out[got:].ctypes.data, out.shape[0] - got, C.byref(em))
            assert rc == 0, m.last_error()
            got += em.value
        L.melspec_stream_destroy(s)
        assert got == want_frames
        outs.append(out[:got].copy())
    batch = h.compute_mel_spectrogram(x[80:])[:want_frames]
    assert np.abs(outs[0] - batch).max() <= 5e-5
    assert np.abs(outs[1] - batch).max() <= 5e-5
    ref = o.whisper_mel_stream(x[:160 * 2000], 400, 160, 80, 16000.0)
    assert np.abs(outs[0][:ref.shape[0]] - ref).max() <= WHISPER_TOL
    h.close()


# ------------------------------------------------------------------------------------------ NeMo BatchLogMel
# The reference computes this path in f32 and holds no output golden for it (parity unpinned beyond the filterbank and the
# shape, see oracle/melspec_oracle.py).  Bar: against the f64 restatement of its arithmetic, max-abs <= 5e-3 and >= 99.5 %
# within 1e-3 for the raw log-mel features (the reference's own f32 pipeline is 3e-4 max away from f64 on JFK).
def _nemo_check(got, want, tol_max=5e-3):
    assert got.shape == want.shape
    d = np.abs(got - want)
    assert d.max() <= tol_max, d.max()
    assert (d <= 1e-3).mean() >= 0.995


def test_nemo_shape_contract(m):
    # src/mel.rs:943-961: 16000 zeros, 128 mels, preemphasis 0.97, guard 2^-24, per-feature normalisation => (128, 101)
    cfg = m.BatchLogMelConfig(n_mels=128, preemphasis=0.97, log_zero_guard=2.0 ** -24, normalize_per_feature=True)
    fe = m.BatchLogMelSpectrogram(cfg)
    f = fe.compute(np.zeros(16000, np.float32))
    assert f.shape == (128, 101) and np.isfinite(f).all()
    assert fe.compute(np.zeros(0, np.float32)).shape == (128, 0)
    flat = fe.compute_flat(np.zeros(16000, np.float32))
    assert (flat.rows, flat.cols, flat.data.size) == (128, 101, 128 * 101)
    with pytest.raises(m.BatchLogMelError):
        m.BatchLogMelSpectrogram(m.BatchLogMelConfig(win_length=600))
    fe.close()


@pytest.mark.parametrize("n_mels,pre,pad_to", [(80, 0.0, 0), (128, 0.97, 0), (128, 0.97, 16)])
def test_nemo_features_vs_oracle(m, jfk, n_mels, pre, pad_to):
    guard = 2.0 ** -24
    fe = m.BatchLogMelSpectrogram(m.BatchLogMelConfig(n_mels=n_mels, preemphasis=pre, log_zero_guard=guard, pad_to=pad_to))
    for x in (jfk, jfk[:16003], o.synth_clip(3, 48000), o.synth_clip(7, 16000)[:159], jfk[:1]):
        got = fe.compute(x)
        want = o.batch_log_mel(x, n_mels=n_mels, preemphasis=pre, log_zero_guard=guard, pad_to=pad_to)
        _nemo_check(got, want)
        valid = x.size // 160 + 1
        assert np.all(got[:, valid:] == 0.0)          # pad_to columns are zeros (src/mel.rs:336)
    fe.close()


def test_nemo_per_feature_normalisation(m, jfk):
    guard = 2.0 ** -24
    cfg = dict(n_mels=128, preemphasis=0.97, log_zero_guard=guard, normalize_per_feature=True, pad_to=8)
    fe = m.BatchLogMelSpectrogram(m.BatchLogMelConfig(**cfg))
    got = fe.compute(jfk)
    want = o.batch_log_mel(jfk, **cfg)
    assert got.shape == want.shape == (128, 1104)
    d = np.abs(got - want)
    assert d.max() <= 5e-3 and (d <= 1e-3).mean() >= 0.995
    v = got[:, :1101]
    assert np.abs(v.mean(axis=1)).max() < 1e-4 and np.abs(v.std(axis=1, ddof=1) - 1.0).max() < 1e-3
    fe.close()


def test_long_single_clip_host_call_is_pipelined_and_identical(m, torch):
    """`compute_mel_spectrogram(&[f32])` on several minutes of audio (src/cuda.rs:88-101): the host call cuts one long clip
    along time into pieces of whole warp tiles and pipelines H2D / kernel / D2H; the frames must equal the unsplit device
    launch bit for bit, in both layouts, for a length that is not a multiple of anything."""
    n = 16000 * 60 * 6 + 1234                                   # 6 minutes: 23 MB of PCM -> three 8 MB pieces
    x = np.concatenate([o.synth_clip(i, 160000) for i in range(37)])[:n]
    for fft, hop in ((400, 160), (512, 160), (1024, 256)):
        h = m.CudaMelSpectrogram(fft, hop, 16000.0, 80)
        xd = torch.from_numpy(x).cuda()
        f = h.num_frames(n)
        for layout in (0, 1):
            shape = (f, 80) if layout == 0 else (80, f)
            dev = torch.empty(shape, dtype=torch.float32, device="cuda")
            h.compute_device(xd, 1, n, n, dev, layout=layout)
            torch.cuda.synchronize()
            host = h.compute_host(x, layout=layout)
            assert host.shape == shape
            assert np.array_equal(host, dev.cpu().numpy()), (fft, hop, layout)
        want = o.whisper_mel_batch(x[:160 * 3000 + fft], fft, hop, 80, 16000.0)
        got = h.compute_host(x)[:want.shape[0]]
        assert np.abs(got - want).max() <= WHISPER_TOL
        h.close()
