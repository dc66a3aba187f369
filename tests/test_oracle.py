"""Pins the CPU oracle (oracle/) against every golden vector the reference's own tests hold for the hot path.
CPU only.  Reference tests restated: src/rb.rs:134-179, src/mel.rs:786-871, 887-911, src/stft.rs:175-194,
src/fbank.rs:354-386, 439-535, tests/readme_examples.rs:11-52."""
import os

import numpy as np
import pytest

import melspec_oracle as o
import oracle_c as oc


def test_stream_path_matches_rust_golden(jfk, golden_dir):
    # src/rb.rs:134-179: fft 512 / hop 160 / 80 mels, |d| <= 1e-6 per element
    gold = np.load(os.path.join(golden_dir, "rust_jfk_golden.npy"))
    got = o.whisper_mel_stream(jfk, 512, 160, 80, 16000.0)
    assert got.T.shape == gold.shape == (80, 1097)
    assert np.abs(got.T - gold).max() <= 1e-6


def test_batch_path_is_stream_path_shifted(jfk, golden_dir):
    gold = np.load(os.path.join(golden_dir, "rust_jfk_golden.npy"))
    c = o.stream_offset(512, 160)
    assert c == 128 and o.stream_offset(400, 160) == 80
    got = o.whisper_mel_batch(jfk[c:], 512, 160, 80, 16000.0)
    assert np.abs(got.T - gold).max() <= 1e-6


def test_c_oracle_matches_rust_golden(jfk, golden_dir):
    gold = np.load(os.path.join(golden_dir, "rust_jfk_golden.npy"))
    got = oc.whisper_batch(jfk[128:], 512, 160, 80, 16000.0)[0]
    assert np.abs(got.T - gold).max() <= 1e-6


@pytest.mark.parametrize("fft", [400, 512])
def test_c_oracle_matches_numpy_oracle_whisper(jfk, fft):
    a = oc.whisper_batch(jfk, fft, 160, 80, 16000.0)[0]
    b = o.whisper_mel_batch(jfk, fft, 160, 80, 16000.0)
    assert a.shape == b.shape
    assert np.abs(a - b).max() <= 1e-6


def test_c_oracle_threads_agree():
    x = np.stack([o.synth_clip(i, 16000) for i in range(5)])
    assert np.array_equal(oc.whisper_batch(x, threads=1), oc.whisper_batch(x, threads=3))


def test_mel_filters_golden(golden_dir):
    # src/mel.rs:837-850 and 852-871: <= 1e-7
    assert np.abs(o.slaney_mel_filterbank(16000, 400, 80) - np.load(os.path.join(golden_dir, "mel_filters_80x201.npy"))).max() <= 1e-7
    assert np.abs(o.slaney_mel_filterbank(16000, 512, 80) - np.load(os.path.join(golden_dir, "nemo_filters_80x257.npy"))).max() <= 1e-7
    assert np.abs(oc.slaney_filterbank(16000, 400, 80) - o.slaney_mel_filterbank(16000, 400, 80)).max() <= 1e-12
    assert np.abs(oc.kaldi_filterbank() - o.kaldi_mel_filterbank()).max() <= 1e-12


def test_mel_scale_known_answers():
    # src/mel.rs:786-804
    assert abs(o.hz_to_mel(60.0) - 0.9) < 1e-10
    assert abs(o.mel_to_hz(3.0) - 200.0) < 1e-10
    # librosa mel_frequencies(n_mels=40) end points (src/mel.rs:806-822)
    mf = o.mel_frequencies(40, 0.0, 11025.0)
    assert mf[0] == 0.0 and abs(mf[-1] - 11025.0) < 5e-3 and abs(mf[1] - 85.317) < 5e-3
    ff = o.fft_frequencies(22050.0, 16)    # src/mel.rs:824-835
    assert np.allclose(ff, [0.0, 1378.125, 2756.25, 4134.375, 5512.5, 6890.625, 8268.75, 9646.875, 11025.0])


def test_sparse_structure():
    # src/mel.rs:887-911: sparse == dense, nnz < 10 %
    f = o.slaney_mel_filterbank(16000, 400, 80)
    rows = o.sparse_rows(f)
    nnz = sum(len(b) for b, _ in rows)
    assert nnz * 10 < f.size
    rng = np.random.default_rng(0)
    p = rng.random(201)
    dense = f @ p
    sparse = np.array([np.dot(w, p[b]) for b, w in rows])
    assert np.abs(dense - sparse).max() <= 1e-12
    assert np.all(f[:, 0] == 0.0)             # DC column is zero (the CUDA kernel relies on it)
    assert np.all(o.kaldi_mel_filterbank()[:, 0] == 0.0)


def test_framing_rules():
    # src/stft.rs:153-157, src/fbank.rs:147-151
    assert o.num_frames(399, 400, 160) == 0
    assert o.num_frames(400, 400, 160) == 1
    assert o.num_frames(160000, 400, 160) == 998
    assert o.num_frames(480000, 400, 160) == 2998
    assert o.num_frames(176000, 512, 160) == 1097
    assert o.whisper_mel_batch(np.zeros(100, np.float32)).shape == (0, 80)
    # silence hits the 1e-10 floor: (-10 + 4)/4 = -1.5 everywhere
    assert np.all(o.whisper_mel_batch(np.zeros(16000, np.float32)) == -1.5)


def test_stream_gating():
    # src/stft.rs:175-194 restated on the stream oracle: no frame until fft_size samples were seen
    x = np.ones(160 * 2, np.float32)
    assert o.whisper_mel_stream(x, 400, 160, 80).shape[0] == 0
    assert o.whisper_mel_stream(np.ones(160 * 3, np.float32), 400, 160, 80).shape[0] == 1


def test_kaldi_shape_and_distance_to_knf_golden(jfk, golden_dir):
    # src/fbank.rs:439-535 asserts shape + finiteness + variance only; the distance is the reference's own
    gold = np.load(os.path.join(golden_dir, "kaldi_fbank_jfk.npy")).T
    got = o.kaldi_fbank(jfk)
    assert got.shape == gold.shape == (1098, 80)
    assert np.isfinite(got).all() and got.var() > 0.1
    d = np.abs(got - gold)
    assert d.max() < 2e-2 and d.mean() < 4e-3
    c = oc.kaldi_batch(jfk)[0]
    assert np.abs(c - got).max() < 1e-4      # f32 CMN mean order differs (f64 accumulate vs pairwise f32)
    assert np.abs(oc.kaldi_batch(jfk, cmn=False)[0] - o.kaldi_fbank(jfk, apply_cmn=False)).max() <= 1e-6


def test_kaldi_config_known_answers():
    # src/fbank.rs:354-386
    assert abs(o.kaldi_mel_to_hz(o.kaldi_hz_to_mel(1000.0)) - 1000.0) < 1e-9
    assert o.kaldi_fbank(np.zeros(16000, np.float32)).shape == (98, 80)
    assert o.kaldi_fbank(np.zeros(399, np.float32)).shape == (0, 80)


# ------------------------------------------------------------------------------------------ NeMo BatchLogMel
def test_nemo_oracle_shape_contract_and_framing():
    # src/mel.rs:943-961: 16000 zeros, 128 mels, pre-emphasis 0.97, guard 2^-24, per-feature normalisation => (128, 101)
    f = o.batch_log_mel(np.zeros(16000, np.float32), n_mels=128, preemphasis=0.97, log_zero_guard=2.0 ** -24,
                        normalize_per_feature=True)
    assert f.shape == (128, 101) and np.isfinite(f).all()
    assert o.batch_log_mel(np.zeros(0, np.float32)).shape == (80, 0)
    # src/mel.rs:387-395, 751-756
    assert o.batch_num_frames(16000, 512, 160, True) == 101 and o.batch_num_frames(16000, 512, 160, False) == 97
    assert o.batch_num_frames(511, 512, 160, False) == 0 and o.batch_num_frames(1, 512, 160, True) == 1
    assert o.pad_len(101, 0) == 101 and o.pad_len(101, 16) == 112 and o.pad_len(112, 16) == 112
    # silence: every valid column is ln(guard); pad_to columns stay zero (src/mel.rs:336)
    g = o.batch_log_mel(np.zeros(1600, np.float32), log_zero_guard=2.0 ** -24, pad_to=16)
    assert g.shape == (80, 16) and np.allclose(g[:, :11], np.log(2.0 ** -24)) and np.all(g[:, 11:] == 0.0)


def test_nemo_oracle_filterbank_and_window(golden_dir):
    # src/mel.rs:852-871: slaney_mel_filterbank(16000, 512, 80) vs nemo_mel_filters.npz within 1e-7
    gold = np.load(os.path.join(golden_dir, "nemo_filters_80x257.npy"))
    assert np.abs(o.general_mel_filterbank(16000.0, 512, 80) - gold).max() <= 1e-7
    w = o.centered_hann_window(512, 400)
    assert np.all(w[:56] == 0.0) and np.all(w[456:] == 0.0) and abs(w[56 + 199] - w[56 + 200]) < 1e-12
    assert w[56] == 0.0 and abs(w.max() - 1.0) < 1e-4


def test_nemo_oracle_f32_pipeline_distance(jfk):
    # the reference runs this path in f32 (src/mel.rs:321-385); the f64 restatement is within the f32 pipeline's noise
    a = o.batch_log_mel(jfk[:48000], n_mels=128, preemphasis=0.97, log_zero_guard=2.0 ** -24)
    b = o.batch_log_mel(jfk[:48000], n_mels=128, preemphasis=0.97, log_zero_guard=2.0 ** -24, dtype=np.float32)
    assert a.shape == b.shape == (128, 301)
    assert np.abs(a - b).max() < 5e-3


@pytest.mark.parametrize("fft,hop,n_mels", [(1024, 256, 128), (480, 160, 80), (441, 147, 64), (251, 100, 40), (256, 64, 40)])
def test_c_oracle_matches_numpy_oracle_at_general_sizes(fft, hop, n_mels):
    """The two restatements of src/stft.rs:89-169 + src/mel.rs agree at the sizes the general plan is checked on (the C
    one is the timed CPU baseline, the numpy one the parity checker): own mixed-radix FFT vs numpy's pocketfft."""
    import oracle_c as oc
    oc.build()
    x = np.stack([o.synth_clip(i, 30000) for i in range(2)])
    a = oc.whisper_batch(x, fft, hop, n_mels, 16000.0, threads=2)
    b = np.stack([o.whisper_mel_batch(x[i], fft, hop, n_mels, 16000.0) for i in range(2)])
    assert a.shape == b.shape and np.abs(a - b).max() <= 1e-6


def test_spectrogram_add_contract(jfk):
    """src/stft.rs:175-194 (`test_spectrogram_add`, fft 8 / hop 4): 3 samples -> None, 4 more (7 < 8) -> None, 4 more -> Some;
    and fed whole hops the per-call restatement equals the stream restatement that is pinned on rust_jfk_golden.npy."""
    got = o.spectrogram_add_mel([[1.0, 2.0, 3.0], [1.0, 2.0, 3.0, 4.0], [1.0, 2.0, 3.0, 4.0]], 8, 4, 4, 16000.0)
    assert got[0] is None and got[1] is None and got[2] is not None and got[2].shape == (4,)
    x = jfk[:16000]
    hops = [x[i:i + 160] for i in range(0, 16000, 160)]
    per_call = [f for f in o.spectrogram_add_mel(hops, 400, 160, 80, 16000.0) if f is not None]
    assert np.array_equal(np.stack(per_call), o.whisper_mel_stream(x, 400, 160, 80, 16000.0))
