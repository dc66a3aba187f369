"""Boundary cases added in round 2, through the C ABI on a B200 (`pytest -m gpu`):
Spectrogram::add's own contract (src/stft.rs:48-86, test at 175-194), the int16 host entry, host == device for the Kaldi
frontend (fused CMN in every pipeline chunk), buffer sizing of the host call for a padded NeMo frontend, and the lifetime of
the thread-local error string."""
import ctypes as C
import threading

import numpy as np
import pytest

import melspec_oracle as o

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


def test_spectrogram_add_reference_test(m):
    # src/stft.rs:175-194 restated: fft 8, hop 4 -> None, None (7 < 8 true samples), Some
    sp = m.Spectrogram(8, 4, n_mels=4, sampling_rate=16000.0)
    mel = m.MelSpectrogram(8, 16000.0, 4)
    assert sp.add([1.0, 2.0, 3.0]) is None
    assert sp.add([1.0, 2.0, 3.0, 4.0]) is None
    fr = sp.add([1.0, 2.0, 3.0, 4.0])
    assert fr is not None
    want = o.spectrogram_add_mel([[1.0, 2.0, 3.0], [1.0, 2.0, 3.0, 4.0], [1.0, 2.0, 3.0, 4.0]], 8, 4, 4, 16000.0)[2]
    got = mel.add(fr)
    assert got.shape == (4, 1) and np.abs(got[:, 0] - want).max() <= WHISPER_TOL
    with pytest.raises(AssertionError):
        sp.add(np.zeros(5, np.float32))                      # "frames must be <= hop_size", src/stft.rs:53
    sp.close()


@pytest.mark.parametrize("fft,hop", [(400, 160), (512, 160)])
def test_spectrogram_add_short_chunks_match_oracle(m, jfk, fft, hop):
    """Ragged chunk lengths (0..hop samples per call): zero padding of short chunks, idx counting true samples only, a frame
    with every call once idx >= fft_size -- call by call against the restatement of src/stft.rs:48-86."""
    rng = np.random.default_rng(3)
    x = jfk[20000:36000]
    chunks, pos = [], 0
    while pos < x.size and len(chunks) < 160:
        n = int(rng.choice([hop, hop, hop, hop - 1, 97, 1, 0, hop // 2]))
        chunks.append(x[pos:pos + n])
        pos += n
    want = o.spectrogram_add_mel(chunks, fft, hop, 80, 16000.0)
    sp = m.Spectrogram(fft, hop)
    mel = m.MelSpectrogram(fft, 16000.0, 80)
    n_some = 0
    for c, w in zip(chunks, want):
        fr = sp.add(c)
        assert (fr is None) == (w is None)
        if fr is not None:
            n_some += 1
            assert np.abs(mel.add(fr)[:, 0] - w).max() <= WHISPER_TOL
    assert n_some > 100
    sp.close()


def test_int16_host_call_equals_f32_call_bit_for_bit(m, jfk):
    h = m.CudaMelSpectrogram(400, 160, 16000.0, 80)
    rng = np.random.default_rng(0)
    pcm16 = np.stack([np.clip(np.round(jfk[:48000] * 32767.0), -32768, 32767).astype(np.int16),
                      rng.integers(-32768, 32768, 48000, dtype=np.int16),
                      np.zeros(48000, np.int16)])
    got = h.compute_host_i16(pcm16)
    ref = h.compute_host(pcm16.astype(np.float32) / 32768.0)
    assert got.shape == ref.shape == (3, 298, 80)
    assert np.array_equal(got, ref)
    assert np.abs(got[0] - o.whisper_mel_batch(pcm16[0].astype(np.float32) / 32768.0)).max() <= WHISPER_TOL
    assert np.all(got[2] == -1.5)
    # odd length (rows not a multiple of 8 samples), single clip, mel-major
    a = h.compute_host_i16(pcm16[1, :12345], layout=m.LAYOUT_MEL_MAJOR)
    b = h.compute_host(pcm16[1, :12345].astype(np.float32) / 32768.0, layout=m.LAYOUT_MEL_MAJOR)
    assert a.shape == (80, 75) and np.array_equal(a, b)
    assert h.compute_host_i16(np.zeros(10, np.int16)).shape == (0, 80)
    h.close()


def test_int16_long_single_clip_is_pipelined_and_identical(m):
    h = m.CudaMelSpectrogram(400, 160, 16000.0, 80)
    rng = np.random.default_rng(1)
    x16 = rng.integers(-20000, 20000, 16000 * 600, dtype=np.int16)      # 10 minutes: takes the time-split pipeline
    a = h.compute_host_i16(x16)
    b = h.compute_host(x16.astype(np.float32) / 32768.0)
    assert a.shape == (59998, 80) and np.array_equal(a, b)
    h.close()


def test_kaldi_host_batch_matches_device_batch(m, torch):
    """>= one clip per SM with CMN: every pipeline chunk of the host call takes the fused-CMN kernel, as the device-resident
    launch of the whole batch does, so the two are bit-identical (and a batch size that does not divide evenly leaves no
    runt chunk below the fused threshold)."""
    fb = m.Fbank(m.FbankConfig())
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    for clips in (2 * sms + 5, 3 * sms + 1):
        n = 16000 * 4
        g = torch.Generator(device="cuda").manual_seed(clips)
        x = (0.1 * torch.randn((clips, n), device="cuda", generator=g)).contiguous()
        F = fb.num_frames(n)
        dev = torch.empty((clips, F, 80), dtype=torch.float32, device="cuda")
        fb.compute_device(x, clips, n, n, dev)
        torch.cuda.synchronize()
        host = fb.compute_host(x.cpu().numpy())
        assert np.array_equal(host, dev.cpu().numpy())
        want = o.kaldi_fbank(x[3].cpu().numpy())
        d = np.abs(host[3] - want)
        assert d.max() <= 5e-3 and (d <= 1e-3).mean() >= 0.995
    fb.close()


def test_host_call_sizes_buffer_with_padded_frames(m, jfk):
    """BatchLogMelSpectrogram(pad_to > 0).compute_host must allocate melspec_padded_frames columns per clip."""
    fe = m.BatchLogMelSpectrogram(m.BatchLogMelConfig(n_mels=80, pad_to=16))
    x = np.stack([jfk[:16000], jfk[16000:32000]])
    out = fe.compute_host(x, layout=m.LAYOUT_MEL_MAJOR)
    cols = fe.padded_frames(16000)
    assert cols == 112 and out.shape == (2, 80, cols)
    assert np.array_equal(out[0], fe.compute(x[0])) and np.array_equal(out[1], fe.compute(x[1]))
    with pytest.raises(ValueError):
        fe.compute_host(x, layout=m.LAYOUT_MEL_MAJOR, out=np.zeros((2, 80, 101), np.float32))
    fe.close()


def test_stream_create_rejects_frontends_without_stream_semantics(m):
    L = m.lib()
    for h in (m.Fbank(m.FbankConfig()), m.BatchLogMelSpectrogram(m.BatchLogMelConfig())):
        s = C.c_void_p()
        assert L.melspec_stream_create(h._h, 16000, C.byref(s)) == 5          # MELSPEC_ERR_UNSUPPORTED
        assert "Whisper" in m.last_error()
        h.close()
    g = m.CudaMelSpectrogram(64, 100, 16000.0, 20)                             # hop > fft: the reference's add() panics
    s = C.c_void_p()
    assert L.melspec_stream_create(g._h, 16000, C.byref(s)) == 5
    g.close()


def test_last_error_is_per_thread_and_survives_successful_calls(m):
    L = m.lib()
    h = m.CudaMelSpectrogram(400, 160, 16000.0, 80)
    assert L.melspec_compute_device(h._h, None, 1, 16000, 16000, None, None, 0, 7, None) != 0      # unknown layout
    mine = m.last_error()
    assert "layout" in mine
    seen = {}

    def other():
        seen["before"] = m.last_error()                       # this thread never failed: its own string is empty
        cfg = m.lib().melspec_default_config(9, None)         # fails in THIS thread only
        seen["code"] = cfg
        seen["after"] = m.last_error()

    t = threading.Thread(target=other)
    t.start(); t.join()
    assert seen["before"] == "" and seen["code"] != 0 and seen["after"] != ""
    assert h.num_frames(16000) == 98                          # a successful call does not touch the string
    assert m.last_error() == mine                             # nor does another thread's failure
    h.close()
