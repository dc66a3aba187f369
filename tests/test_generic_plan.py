"""The general plan (mel-spec_b200/csrc/melspec_generic.cuh): every fft_size / hop / frame length / sample rate the
reference accepts (rustfft plans any length; src/stft.rs:119-138, src/fbank.rs:25-82, src/mel.rs:171-214) that the two
specialised kernels do not cover, through the same C ABI, against the f64 oracle.  Needs a B200: `pytest -m gpu`.

Tolerances are the ones of tests/test_gpu_parity.py: Whisper path <= 1e-4 max-abs; ln() paths (Kaldi, NeMo) <= 5e-3 max-abs
with >= 99.5 % of the values within 1e-3 (unclamped logarithms of fp32 energies)."""
import numpy as np
import pytest

import melspec_oracle as o

pytestmark = pytest.mark.gpu

WHISPER_TOL = 1e-4
LN_TOL_MAX = 5e-3
LN_TOL_BULK = 1e-3


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


def _ln_check(got, want):
    assert got.shape == want.shape
    if want.size == 0:
        return
    d = np.abs(got - want)
    assert d.max() <= LN_TOL_MAX, d.max()
    assert (d <= LN_TOL_BULK).mean() >= 0.995, (d <= LN_TOL_BULK).mean()


# fft sizes: powers of two either side of the specialised ones, 2^a 3^b 5^c composites, odd composites, a prime, tiny
WHISPER_CASES = [
    (1024, 256, 128, 16000.0), (256, 64, 40, 8000.0), (2048, 512, 80, 44100.0), (4096, 1024, 128, 48000.0),
    (480, 160, 80, 16000.0), (441, 147, 64, 22050.0), (360, 90, 32, 12000.0), (251, 100, 40, 16000.0),
    (400, 320, 80, 16000.0), (512, 128, 80, 16000.0), (512, 160, 80, 22050.0), (16, 4, 4, 16000.0), (8192, 2048, 80, 48000.0),
]


@pytest.mark.parametrize("fft,hop,n_mels,sr", WHISPER_CASES)
def test_whisper_any_size_vs_oracle(m, fft, hop, n_mels, sr):
    rng = np.random.default_rng(fft * 7 + hop)
    n = max(8 * fft, 20000) + 13
    t = np.arange(n) / sr
    x = (0.5 * np.sin(2 * np.pi * 0.011 * sr * t) + 0.2 * np.sin(2 * np.pi * 0.13 * sr * t) + 0.05 * rng.standard_normal(n)).astype(np.float32)
    x[: n // 5] *= 1e-3                                     # a quiet stretch: exercises the max-8 clamp region
    h = m.CudaMelSpectrogram(fft, hop, sr, n_mels)
    assert h.num_frames(n) == o.num_frames(n, fft, hop)
    got = h.compute_mel_spectrogram(x)
    want = o.whisper_mel_batch(x, fft, hop, n_mels, sr)
    assert got.shape == want.shape and want.shape[0] > 0
    assert np.abs(got - want).max() <= WHISPER_TOL, np.abs(got - want).max()
    # ragged ends and inputs shorter than one frame (src/stft.rs:153-157)
    for cut in (fft - 1, fft, fft + hop - 1, fft + hop, fft + 3 * hop + 1):
        g2 = h.compute_mel_spectrogram(x[:cut])
        w2 = o.whisper_mel_batch(x[:cut], fft, hop, n_mels, sr)
        assert g2.shape == w2.shape
        if w2.size:
            assert np.abs(g2 - w2).max() <= WHISPER_TOL
    h.close()


def test_generic_fft400_agrees_with_specialised_kernel(m, jfk):
    # hop 320 at fft 400 runs on the general plan; its frames are every other frame of the hop-160 kernel
    a = m.CudaMelSpectrogram(400, 160, 16000.0, 80)
    b = m.CudaMelSpectrogram(400, 320, 16000.0, 80)
    fa, fb = a.compute_mel_spectrogram(jfk), b.compute_mel_spectrogram(jfk)
    assert fb.shape == (549, 80)
    assert np.abs(fa[::2] - fb).max() <= 6e-5               # two fp32 FFT schedules, each within 3.4e-5 of the f64 oracle
    a.close(), b.close()


def test_generic_batch_layouts_and_lengths(m, torch):
    fft, hop, n_mels, sr = 1024, 256, 128, 16000.0
    h = m.CudaMelSpectrogram(fft, hop, sr, n_mels)
    s = 30000
    lens = [30000, 0, 1023, 1024, 1279, 1280, 29999, 7777]
    rng = np.random.default_rng(11)
    pcm = (rng.standard_normal((len(lens), s)) * 0.2).astype(np.float32)
    x = torch.from_numpy(pcm).cuda()
    f = h.num_frames(s)
    for layout in (0, 1):
        shape = (len(lens), f, n_mels) if layout == 0 else (len(lens), n_mels, f)
        out = torch.full(shape, float("nan"), dtype=torch.float32, device="cuda")
        h.compute_device(x, len(lens), s, s, out, d_lens=torch.tensor(lens, dtype=torch.int32, device="cuda"), layout=layout)
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        if layout == 1:
            got = got.transpose(0, 2, 1)
        for i, n in enumerate(lens):
            want = o.whisper_mel_batch(pcm[i, :n], fft, hop, n_mels, sr)
            k = want.shape[0]
            if k:
                assert np.abs(got[i, :k] - want).max() <= WHISPER_TOL
            assert np.isnan(got[i, k:]).all(), "frames past a clip's own length must stay untouched"
    h.close()


def test_generic_streaming_ringbuffer(m, jfk):
    # RingBuffer semantics (src/rb.rs:86-121, src/stft.rs:48-86) at a size neither specialised kernel covers
    fft, hop, n_mels = 1024, 256, 80
    rb = m.RingBuffer(m.MelConfig(fft, hop, n_mels, 16000.0), capacity=1 << 20)
    x = jfk[:60000 + 101]
    rng = np.random.default_rng(2)
    pos, frames = 0, []
    while pos < x.size:
        n = int(rng.integers(1, 3000))
        rb.add_frame(x[pos:pos + n])
        pos += n
        while True:
            fr = rb.maybe_mel()
            if fr is None:
                break
            frames.append(fr[:, 0])
    got = np.stack(frames)
    want = o.whisper_mel_stream(x, fft, hop, n_mels, 16000.0)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= WHISPER_TOL
    rb.close()


KALDI_CASES = [
    dict(sample_rate=8000.0, num_mel_bins=40),                                  # 200-sample frames, fft 256, shift 80
    dict(sample_rate=16000.0, frame_length_ms=32.0, frame_shift_ms=16.0),        # 512-sample frames = the fft size
    dict(sample_rate=44100.0, num_mel_bins=64, frame_length_ms=25.0),            # 1102-sample frames, fft 2048, shift 441
    dict(use_power=False),                                                       # magnitude spectrum (src/fbank.rs:197-203)
    dict(use_log_fbank=False, apply_cmn=False),
    dict(preemphasis=0.0, energy_floor=1e-3, low_freq=100.0, high_freq=7000.0, num_mel_bins=23, sample_rate=22050.0),
]


@pytest.mark.parametrize("kw", KALDI_CASES)
def test_kaldi_any_config_vs_oracle(m, jfk, kw):
    fb = m.Fbank(m.FbankConfig(**kw))
    for x in (jfk[:50000], jfk[:50000 - 337], jfk[:150]):
        got = fb.compute(x)
        want = o.kaldi_fbank(x, **kw)
        if kw.get("use_log_fbank", True):
            _ln_check(got, want)
        else:                                                   # linear energies: relative comparison
            assert got.shape == want.shape
            if want.size:
                assert np.abs(got - want).max() <= 1e-4 * max(1.0, np.abs(want).max())
    fb.close()


NEMO_CASES = [
    dict(sample_rate=22050, n_fft=1024, win_length=1024, hop_length=256, n_mels=80),
    dict(n_fft=512, win_length=320, hop_length=160, n_mels=64, preemphasis=0.97),
    dict(n_fft=400, win_length=400, hop_length=160, n_mels=80, center=False, pad_to=16),
    dict(n_fft=1024, win_length=800, hop_length=200, n_mels=128, htk=True, norm=False, f_min=50.0, f_max=7600.0,
         normalize_per_feature=True, preemphasis=0.97),
    dict(n_fft=768, win_length=601, hop_length=123, n_mels=96),                 # odd window, 2^8 * 3
    dict(n_fft=512, win_length=400, hop_length=160, n_mels=40, f_min=-300.0),   # bank with a non-zero DC column -> general plan
]


@pytest.mark.parametrize("kw", NEMO_CASES)
def test_nemo_any_config_vs_oracle(m, jfk, kw):
    guard = 2.0 ** -24
    fe = m.BatchLogMelSpectrogram(m.BatchLogMelConfig(log_zero_guard=guard, **kw))
    for x in (jfk[:40000], jfk[:12345], jfk[:100], jfk[:1]):
        got = fe.compute(x)
        want = o.batch_log_mel(x, log_zero_guard=guard, **kw)
        _ln_check(got, want)
    fe.close()


def test_generic_interleaved_image_and_tga(m, jfk):
    # the formats after the path (src/mel.rs:480-544, src/quant.rs:38-64) on a size that runs on the general plan
    fft, hop, n_mels = 1024, 256, 80
    h = m.CudaMelSpectrogram(fft, hop, 16000.0, n_mels)
    x = jfk[:70000]
    frames = h.compute_mel_spectrogram(x)
    want = o.whisper_mel_batch(x, fft, hop, n_mels, 16000.0)
    assert np.abs(frames - want).max() <= WHISPER_TOL
    f = frames.shape[0]                                        # 270 frames -> even width 270; min_width 300 pads with zeros
    for min_width in (0, 300):
        tga, img = h.mel_tga(x, min_width=min_width, return_image=True)
        w = img.shape[1]
        assert w == o.interleave_frames(want, False, min_width).size // n_mels
        assert np.array_equal(img[:, :f], frames.T), "interleaved image == frames transposed"
        assert np.all(img[:, f:] == 0.0)
        assert tga == o.tga_8bit_data(img.reshape(-1), n_mels), "device quantiser bytes == quant.rs arithmetic on the same f32 image"
    h.close()


def test_reference_goldens_hold_on_the_general_plan():
    """MELSPEC_FORCE_GENERIC=1 routes every configuration through the general kernel: the reference's own fixtures
    (rust_jfk_golden.npy at fft 512, quantized_mel_golden.tga at fft 400, the Kaldi and NeMo contracts, the C++ host mirror's
    restatement of the reference tests) must hold there too — a second, independent implementation of the same transforms.
    Run in a child process because the switch is read once per process."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MELSPEC_FORCE_GENERIC="1")
    sel = "golden or jfk or readme or reference_cuda or nemo_shape or ringbuffer or streaming_matches or cpp_host_mirror_reference"
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_parity.py"),
                          os.path.join(root, "tests", "test_formats.py"), os.path.join(root, "tests", "test_cpp_host.py"),
                          "-m", "gpu", "-q", "-x", "-k", sel], capture_output=True, text=True, env=env, cwd=root, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert " passed" in out.stdout and "failed" not in out.stdout


@pytest.mark.parametrize("kw", [dict(n_mels=80, pad_to=16), dict(n_mels=128, preemphasis=0.97, normalize_per_feature=True, pad_to=8),
                                dict(n_fft=1024, win_length=1024, hop_length=256, n_mels=64, center=False)])
def test_nemo_ragged_batch(m, torch, kw):
    """Per-clip lengths for the NeMo frontend (a batch extension of the reference's single-waveform API): every clip must equal
    `BatchLogMelSpectrogram::compute` of its own samples (src/mel.rs:321-385), zero-padded to the batch's common width."""
    guard = 2.0 ** -24
    fe = m.BatchLogMelSpectrogram(m.BatchLogMelConfig(log_zero_guard=guard, **kw))
    s = 40000
    lens = [40000, 0, 1, 159, 160, 12345, 39999, 20000, 1023, 1024]
    rng = np.random.default_rng(21)
    pcm = (rng.standard_normal((len(lens), s)) * 0.1).astype(np.float32)
    cols = fe.padded_frames(s)
    out = torch.full((len(lens), fe.n_mels, cols), float("nan"), dtype=torch.float32, device="cuda")
    fe.compute_device(torch.from_numpy(pcm).cuda(), len(lens), s, s, out, layout=m.LAYOUT_MEL_MAJOR,
                      d_lens=torch.tensor(lens, dtype=torch.int32, device="cuda"))
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert np.isfinite(got).all()
    for i, n in enumerate(lens):
        want = o.batch_log_mel(pcm[i, :n], log_zero_guard=guard, **kw)
        w = want.shape[1]
        if w:
            d = np.abs(got[i][:, :w] - want)
            assert d.max() <= LN_TOL_MAX and (d <= LN_TOL_BULK).mean() >= 0.995, (i, n, d.max())
        assert np.all(got[i][:, w:] == 0.0), (i, n)
    fe.close()


@pytest.mark.parametrize("fft,hop,n_mels", [(480, 160, 80), (1024, 256, 128), (256, 64, 40), (450, 150, 64)])
def test_pair_form_unaligned_and_odd_counts(m, torch, fft, hop, n_mels):
    """The pair form of the general plan (melspec_generic2.cuh: two frames per warp) on inputs its fast paths must refuse: a PCM
    pointer that is only 4-byte aligned (no 64-bit loads), odd and even frame counts (a last pair with one frame), both layouts."""
    sr = 16000.0
    h = m.CudaMelSpectrogram(fft, hop, sr, n_mels)
    rng = np.random.default_rng(fft + 3)
    for frames in (1, 2, 7, 32, 33):
        n = fft + (frames - 1) * hop + 5
        pcm = (rng.standard_normal((3, n + 1)) * 0.3).astype(np.float32)
        base = torch.from_numpy(pcm).cuda()
        for off in (0, 1):
            flat = base.reshape(-1)[off:]                      # off = 1: every clip starts on an odd word
            stride = n + 1
            ns = n if off == 0 else n - 1                      # stay inside the allocation for the last clip
            assert h.num_frames(ns) in (frames, frames - 1)
            f = h.num_frames(ns)
            if f == 0:
                continue
            for layout in (0, 1):
                shape = (3, f, n_mels) if layout == 0 else (3, n_mels, f)
                out = torch.full(shape, float("nan"), dtype=torch.float32, device="cuda")
                h.compute_device(flat, 3, stride, ns, out, layout=layout)
                torch.cuda.synchronize()
                got = out.cpu().numpy()
                if layout == 1:
                    got = got.transpose(0, 2, 1)
                for i in range(3):
                    x = pcm.reshape(-1)[off + i * stride: off + i * stride + ns]
                    want = o.whisper_mel_batch(x, fft, hop, n_mels, sr)
                    assert want.shape == got[i].shape
                    assert np.abs(got[i] - want).max() <= WHISPER_TOL, (frames, off, layout, i)
    h.close()


def test_pair_form_agrees_with_one_frame_kernel(m, jfk):
    """Both forms of the general plan run the same Stockham schedule per frame; they differ in where the twiddles come from (table
    entries vs products of table entries) and in the untangle step (two bins per step in the pair form), so their outputs agree to
    within twice the distance either keeps from the f64 oracle.  The one-frame
    kernel is selected in a child process (MELSPEC_GENERIC_PAIR is read once per process)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, numpy as np; sys.path.insert(0, %r); import mel_spec_b200 as ms; x = np.load(sys.argv[1]);"
            "h = ms.CudaMelSpectrogram(480, 160, 16000.0, 80); np.save(sys.argv[2], h.compute_mel_spectrogram(x));"
            "k = ms.Fbank(ms.FbankConfig(sample_rate=8000.0, num_mel_bins=40)); np.save(sys.argv[3], k.compute(x))") % root
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        xin = os.path.join(td, "x.npy")
        np.save(xin, jfk[:48000])
        outs = {}
        for mode in ("0", "1"):
            a, b = os.path.join(td, f"w{mode}.npy"), os.path.join(td, f"k{mode}.npy")
            r = subprocess.run([sys.executable, "-c", code, xin, a, b], env=dict(os.environ, MELSPEC_GENERIC_PAIR=mode), cwd=root,
                               capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, r.stderr[-2000:]
            outs[mode] = (np.load(a), np.load(b))
    assert outs["0"][0].shape == outs["1"][0].shape and outs["0"][1].shape == outs["1"][1].shape
    assert np.abs(outs["0"][0] - outs["1"][0]).max() <= 6e-5   # two fp32 schedules, each within ~3e-5 of the f64 oracle
    _ln_check(outs["0"][1], outs["1"][1])
