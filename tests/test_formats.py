"""Output formats: interleave_frames (reference src/mel.rs:480-544) and the 8-bit TGA quantiser
(src/quant.rs:38-165).  CPU tests pin the oracle against the reference's own `testdata/quantized_mel_golden.tga`
(the fixture of src/vad.rs:712,742 and tests/vad_regression.rs:157,215); GPU tests compare the device kernels with the
oracle bit for bit and with the same golden file."""
import os

import numpy as np
import pytest

import melspec_oracle as o


def _golden(golden_dir):
    raw = open(os.path.join(golden_dir, "quantized_mel_golden.tga"), "rb").read()
    assert raw[:18] == bytes([8, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1100 & 255, 1100 >> 8, 80, 0, 8, 0])
    mn, mx = np.frombuffer(raw[18:26], dtype="<f4")
    return raw, np.frombuffer(raw[26:], dtype=np.uint8).reshape(80, 1100), float(mn), float(mx)


# ------------------------------------------------------------------------------------------------ CPU: oracle pinned
def test_oracle_fft400_stream_mel_matches_golden_tga(jfk, golden_dir):
    """The reference's golden TGA is 1100 columns: two leading silent columns (written by an older stream framing) and the
    1098 Whisper fft-400 / hop-160 stream frames of the JFK clip.  Quantising the oracle's frames with the reference's
    arithmetic reproduces every byte AND the f32 min/max of the header: the fft-400 path of the oracle is pinned."""
    raw, img, mn, mx = _golden(golden_dir)
    mel = o.whisper_mel_stream(jfk, 400, 160, 80, 16000.0)            # (1098, 80)
    assert mel.shape == (1098, 80)
    flat = o.interleave_frames(mel, False, 0)
    q, (qmn, qmx) = o.quantize(flat)
    assert (qmn, qmx) == (mn, mx) == (-1.5, float(np.float32(1.535932183265686)))
    assert np.array_equal(q.reshape(80, 1098), img[:, 2:])
    assert not img[:, :2].any()                                         # silence quantises to 0 (= -1.5)


def test_oracle_tga_roundtrip_and_header(jfk):
    mel = o.whisper_mel_batch(jfk[:32000], 400, 160, 80, 16000.0)
    flat = o.interleave_frames(mel, False, 200)
    assert flat.size == 80 * 200 and mel.shape[0] == 198
    assert np.array_equal(flat.reshape(80, 200)[:, :198], mel.T.astype(np.float32)) and not flat.reshape(80, 200)[:, 198:].any()
    tga = o.tga_8bit_data(flat, 80)
    assert len(tga) == 26 + flat.size and tga[12:16] == (200).to_bytes(2, "little") + (80).to_bytes(2, "little")
    back = o.parse_tga_8bit(tga)
    step = (flat.max() - flat.min()) / 255.0
    assert np.abs(back - flat).max() <= 0.5 * step * (1 + 1e-5) + 1e-6    # src/quant.rs: round-to-nearest level


def test_oracle_interleave_rules():
    fr = np.arange(15, dtype=np.float64).reshape(5, 3)                    # 5 frames x 3 mels
    assert o.interleave_frames(fr, False, 0).reshape(3, 5).tolist() == fr.T.tolist()           # no padding when min_width == 0
    w = o.interleave_frames(fr, False, 2).reshape(3, 6)                    # odd count + min_width > 0 -> one zero frame
    assert w[:, :5].tolist() == fr.T.tolist() and not w[:, 5].any()
    w = o.interleave_frames(fr, False, 10).reshape(3, 10)
    assert not w[:, 5:].any()
    c = o.interleave_frames(fr, True, 8)                                   # column-major: frame after frame, then zeros
    assert c.size == 24 and c[:15].tolist() == fr.reshape(-1).tolist() and not c[15:].any()
    with pytest.raises(AssertionError):
        o.interleave_frames(fr, False, 3)
    with pytest.raises(AssertionError):
        o.interleave_frames(np.zeros((0, 3)), False, 0)


def test_oracle_quantize_edge_cases():
    q, r = o.quantize(np.array([0.0, 0.5, 1.0], np.float32))
    assert q.tolist() == [0, 128, 255] and r == (0.0, 1.0)                # 127.5 rounds away from zero
    q, r = o.quantize(np.full(7, 3.25, np.float32))                        # max == min: scale = inf, 0 * inf = NaN -> 0
    assert q.tolist() == [0] * 7 and r == (3.25, 3.25)
    d = o.dequantize(np.array([0, 255], np.uint8), (-1.5, 1.5))
    assert d.tolist() == [-1.5, 1.5]


def test_abi_exports_format_symbols():
    import mel_spec_b200 as ms
    ms.build()
    L = ms.lib()
    assert L.melspec_interleaved_width(1098, 0) == 1098 and L.melspec_interleaved_width(1097, 0) == 1097
    assert L.melspec_interleaved_width(1097, 2) == 1098 and L.melspec_interleaved_width(1097, 3000) == 3000
    assert L.melspec_interleaved_width(10, 3) == -1 and L.melspec_interleaved_width(0, 0) == -1
    assert L.melspec_tga_size(80, 1100) == 88026 and L.melspec_tga_size(80, 65536) == -1 and L.melspec_tga_size(80, 65535) > 0


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def mel400():
    import mel_spec_b200 as ms
    ms.build()
    h = ms.CudaMelSpectrogram(400, 160, 16000.0, 80)
    yield h
    h.close()


@pytest.mark.gpu
def test_gpu_quantize_is_bit_exact_with_oracle(mel400, jfk):
    rng = np.random.default_rng(7)
    mel = o.whisper_mel_batch(jfk[:64000], 400, 160, 80, 16000.0).astype(np.float32)
    for img in (o.interleave_frames(mel, False, 0), o.interleave_frames(mel, False, 402),
                rng.standard_normal(80 * 37).astype(np.float32), rng.standard_normal(3 * 1001).astype(np.float32) * 1e-3,
                np.full(80 * 4, 2.5, np.float32)):
        n_mels = 3 if img.size == 3003 else 80
        assert mel400.tga_8bit_data(img, n_mels) == o.tga_8bit_data(img, n_mels)
        tga = o.tga_8bit_data(img, n_mels)
        assert np.array_equal(mel400.parse_tga_8bit(tga), o.parse_tga_8bit(tga))
    q, r = mel400.quantize(np.array([0.0, 0.5, 1.0], np.float32))
    assert q.tolist() == [0, 128, 255] and (r.min, r.max) == (0.0, 1.0)


@pytest.mark.gpu
def test_gpu_mel_tga_matches_reference_golden(mel400, jfk, golden_dir):
    """PCM -> fused kernel (mel-major store = interleave_frames) -> device quantiser, against the reference's golden TGA.
    fp32 kernel vs f64 reference: a value within 3.4e-5 of a rounding boundary may land on the neighbouring level."""
    raw, img, mn, mx = _golden(golden_dir)
    tga, f32img = mel400.mel_tga(jfk[80:], 0, return_image=True)          # stream framing = batch framing on samples[80..]
    assert f32img.shape == (80, 1098) and len(tga) == 26 + 80 * 1098
    want = o.whisper_mel_batch(jfk[80:], 400, 160, 80, 16000.0).T
    assert np.abs(f32img - want).max() <= 1e-4
    gmn, gmx = np.frombuffer(tga[18:26], dtype="<f4")
    assert gmn == -1.5 and abs(float(gmx) - mx) <= 1e-4
    got = np.frombuffer(tga[26:], dtype=np.uint8).reshape(80, 1098).astype(np.int32)
    diff = np.abs(got - img[:, 2:].astype(np.int32))
    assert diff.max() <= 1 and (diff != 0).mean() <= 0.01, (diff.max(), (diff != 0).mean())
    # and bit-exact with the oracle's quantiser applied to the kernel's own f32 image
    assert tga == o.tga_8bit_data(f32img.reshape(-1), 80)


@pytest.mark.gpu
def test_gpu_interleaved_device_batch(mel400):
    import torch
    pcm = np.stack([o.synth_clip(i, 16000 + 160) for i in range(5)]).astype(np.float32)   # 99 frames: odd
    x = torch.from_numpy(pcm).cuda()
    f = mel400.num_frames(pcm.shape[1])
    assert f == 99
    for min_width, w in ((0, 99), (2, 100), (128, 128)):
        out = torch.full((5, 80, w), float("nan"), dtype=torch.float32, device="cuda")
        mel400.compute_interleaved_device(x, 5, pcm.shape[1], pcm.shape[1], min_width, out)
        tga = torch.zeros((5, 26 + 80 * w), dtype=torch.uint8, device="cuda")
        mel400.quantize_tga_device(out, 5, 80, w, tga)
        back = torch.empty((5, 80, w), dtype=torch.float32, device="cuda")
        mel400.dequantize_tga_device(tga, 5, 80, w, back)
        torch.cuda.synchronize()
        got, tg, bk = out.cpu().numpy(), tga.cpu().numpy(), back.cpu().numpy()
        for i in range(5):
            want = o.interleave_frames(o.whisper_mel_batch(pcm[i], 400, 160, 80, 16000.0), False, min_width).reshape(80, w)
            assert np.abs(got[i] - want).max() <= 1e-4
            assert not got[i][:, f:].any()
            assert tg[i].tobytes() == o.tga_8bit_data(got[i].reshape(-1), 80)
            assert np.array_equal(bk[i].reshape(-1), o.parse_tga_8bit(tg[i].tobytes()))
    assert np.allclose(mel400.interleave_frames(pcm[0], False, 128).reshape(80, 128), got[0], atol=0, rtol=0)
    col = mel400.interleave_frames(pcm[0], True, 128)
    assert col.size == 80 * 128 and np.array_equal(col[:99 * 80].reshape(99, 80), got[0][:, :99].T) and not col[99 * 80:].any()


@pytest.mark.gpu
def test_gpu_format_errors(mel400, jfk):
    with pytest.raises(ValueError):
        mel400.mel_tga(jfk[:16000], 3)                                     # odd min_width (src/mel.rs:488)
    with pytest.raises(ValueError):
        mel400.mel_tga(jfk[:100], 0)                                       # no frame (src/mel.rs:487)
    with pytest.raises(ValueError):
        mel400.tga_8bit_data(np.zeros(80 * 65536, np.float32), 80)         # width does not fit the u16 header field
    with pytest.raises(IOError):
        mel400.parse_tga_8bit(b"\0" * 10)


@pytest.mark.gpu
def test_tga_8bit_chunks_wide_images_like_the_reference(mel400, tmp_path):
    """src/quant.rs:29-36, 100-137: strides of u16::MAX columns, one TGA (own min / max) per stride; save_tga_8bit asserts
    width < u16::MAX (src/quant.rs:17-21)."""
    rng = np.random.default_rng(5)
    n_mels, width = 4, 65535 + 1000
    img = (rng.standard_normal((n_mels, width)) * 0.5).astype(np.float32)
    img[:, 65535:] += 3.0                                                  # the second stride has its own range
    chunks = mel400.tga_8bit(img.reshape(-1), n_mels)
    assert len(chunks) == 2
    for blob, blk in zip(chunks, (img[:, :65535], img[:, 65535:])):
        want = o.tga_8bit_data(np.ascontiguousarray(blk).reshape(-1), n_mels)
        assert blob == want
    assert len(mel400.tga_8bit(img[:, :65535].reshape(-1), n_mels)) == 1   # stride_size == width: a single image
    with pytest.raises(AssertionError):
        mel400.save_tga_8bit(img[:, :65535].reshape(-1), n_mels, str(tmp_path / "x.tga"))
    mel400.save_tga_8bit(img[:, :2000].reshape(-1), n_mels, str(tmp_path / "y.tga"))
    back = mel400.load_tga_8bit(str(tmp_path / "y.tga"))
    assert back.shape == (n_mels * 2000,) and np.abs(back - img[:, :2000].reshape(-1)).max() <= (img[:, :2000].max() - img[:, :2000].min()) / 255.0


@pytest.mark.gpu
def test_gpu_mel_tga_batch_equals_single_clip_calls(mel400, jfk):
    """`melspec_mel_tga_host_batch` (pipelined PCM -> mel -> interleave -> 8-bit TGA for a batch; one byte per mel value crosses
    PCIe on the way back) gives every clip exactly the bytes of the single-clip call, for f32 and int16 PCM, even and padded widths
    and a strided output; the single-clip call is pinned on the reference's quantized_mel_golden.tga above."""
    n = 16000 * 3 + 160                                               # 299 frames: odd, so min_width > 0 adds the zero frame
    clips = np.stack([jfk[i * 7000: i * 7000 + n] * s for i, s in enumerate((1.0, 0.3, 1e-3, 0.0, 2.5))]).astype(np.float32)
    clips[3, 100] = 1e-4                                              # a single tick in an otherwise silent clip
    for min_width in (0, 300, 400):
        got = mel400.mel_tga_batch(clips, min_width=min_width)
        for i in range(clips.shape[0]):
            assert got[i] == mel400.mel_tga(clips[i], min_width=min_width), (min_width, i)
    # int16 PCM: identical to the f32 call on x / 32768
    x16 = np.clip(np.round(clips * 20000.0), -32768, 32767).astype(np.int16)
    got16 = mel400.mel_tga_batch(x16)
    for i in range(clips.shape[0]):
        assert got16[i] == mel400.mel_tga(x16[i].astype(np.float32) / 32768.0), i
    # strided output rows + many clips (several pipeline chunks)
    many = np.ascontiguousarray(np.tile(clips, (40, 1)))              # 200 clips
    w = mel400.interleaved_width(n)
    size = int(mel400._L.melspec_tga_size(mel400.n_mels, w))
    stride = size + 37
    buf = np.full((many.shape[0], stride), 0xAB, dtype=np.uint8)
    wout = mel400.mel_tga_batch_raw(many.ctypes.data, many.shape[0], n, n, buf.ctypes.data, tga_stride=stride)
    assert wout == w
    single = mel400.mel_tga_batch(clips)
    for i in range(many.shape[0]):
        assert buf[i, :size].tobytes() == single[i % 5], i
        assert np.all(buf[i, size:] == 0xAB), "bytes between the images stay untouched"
    # errors: odd min_width, empty frames, width beyond the TGA header's u16
    with pytest.raises(ValueError):
        mel400.mel_tga_batch(clips, min_width=3)
    with pytest.raises(ValueError):
        mel400.mel_tga_batch(clips[:, :100])
