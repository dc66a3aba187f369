"""Pin of the NeMo / Parakeet frontend oracle (reference src/mel.rs:171-418, 656-756).  CPU only.

The reference holds no output fixture for `BatchLogMelSpectrogram`; what it publishes is its distance to NeMo's own
featurizer (README.md:146-158: MAE 0.001183, RMSE 0.0237, max 3.97, correlation 0.999719 on JFK; CHANGELOG.md:42-45).  So the
oracle is pinned on that algorithm: an independent restatement of NeMo's `FilterbankFeatures.forward` on `torch.stft`
(`nemo/collections/asr/parts/preprocessing/features.py`: pre-emphasis of the waveform, `torch.stft(center=True,
pad_mode="constant")` with a symmetric Hann window of win_length, power spectrum, Slaney mel bank, `log(x + guard)`,
per-feature normalisation over seq_len = len // hop + 1 frames with the N-1 variance and +1e-5, zero padding to pad_to).
Nothing here shares code with oracle/melspec_oracle.py except the filterbank, which is itself pinned on the reference's
`nemo_mel_filters.npz` (tests/test_oracle.py)."""
import numpy as np
import pytest
import torch

import melspec_oracle as o


def nemo_filterbank_features(x, n_mels=80, preemph=0.97, guard=2.0 ** -24, normalize=None, pad_to=0, n_fft=512, win=400, hop=160,
                             sr=16000):
    x = torch.from_numpy(np.asarray(x, dtype=np.float32)).to(torch.float64)[None, :]
    seq_len = x.shape[1] // hop + 1
    if preemph:
        x = torch.cat((x[:, :1], x[:, 1:] - preemph * x[:, :-1]), dim=1)
    window = torch.hann_window(win, periodic=False, dtype=torch.float64)
    spec = torch.stft(x, n_fft=n_fft, hop_length=hop, win_length=win, center=True, window=window, return_complex=True,
                      pad_mode="constant")
    power = spec.real ** 2 + spec.imag ** 2                                  # mag_power = 2.0
    fb = torch.from_numpy(o.slaney_mel_filterbank(float(sr), n_fft, n_mels)).to(torch.float64)
    feats = torch.log(torch.matmul(fb, power[0]) + guard)                   # (n_mels, frames)
    feats = feats[:, :seq_len]
    if normalize == "per_feature":
        mean = feats.sum(dim=1, keepdim=True) / seq_len
        std = torch.sqrt(((feats - mean) ** 2).sum(dim=1, keepdim=True) / (seq_len - 1)) + 1e-5
        feats = (feats - mean) / std
    if pad_to:
        cols = -(-seq_len // pad_to) * pad_to
        feats = torch.nn.functional.pad(feats, (0, cols - seq_len), value=0.0)
    return feats.numpy()


@pytest.mark.parametrize("n_mels,preemph", [(80, 0.0), (80, 0.97), (128, 0.97)])
def test_oracle_equals_torch_stft_restatement_of_nemo(jfk, n_mels, preemph):
    guard = 2.0 ** -24
    for x in (jfk, jfk[:16003], o.synth_clip(3, 48000), jfk[:159]):
        want = nemo_filterbank_features(x, n_mels, preemph, guard)
        got = o.batch_log_mel(x, n_mels=n_mels, preemphasis=preemph, log_zero_guard=guard, out_dtype=np.float64)
        assert got.shape == want.shape == (n_mels, x.size // 160 + 1)
        assert np.abs(got - want).max() <= 1e-9, np.abs(got - want).max()


def test_oracle_normalised_and_padded_equals_torch_restatement(jfk):
    guard = 2.0 ** -24
    want = nemo_filterbank_features(jfk, 128, 0.97, guard, normalize="per_feature", pad_to=16)
    raw = o.batch_log_mel(jfk, n_mels=128, preemphasis=0.97, log_zero_guard=guard, out_dtype=np.float64)
    # the oracle normalises the f32-rounded features like the reference (src/mel.rs:721-749 works on the f32 matrix): compare the
    # f64 chain through the same formula, and the f32 product at f32 resolution
    mean = raw.sum(axis=1, keepdims=True) / raw.shape[1]
    std = np.sqrt(((raw - mean) ** 2).sum(axis=1, keepdims=True) / (raw.shape[1] - 1)) + 1e-5
    assert np.abs((raw - mean) / std - want[:, :raw.shape[1]]).max() <= 1e-9
    got = o.batch_log_mel(jfk, n_mels=128, preemphasis=0.97, log_zero_guard=guard, normalize_per_feature=True, pad_to=16)
    assert got.shape == want.shape == (128, 1104)
    assert np.abs(got - want).max() <= 2e-6
    assert np.all(got[:, 1101:] == 0.0)


def test_reference_published_distance_to_nemo_holds_for_the_f32_rerun(jfk):
    """README.md:153-158: the reference's f32 pipeline vs NeMo on JFK: MAE 0.001183, correlation 0.999719 (with per-feature
    normalisation, the Parakeet configuration).  The f32 rerun of the restatement (f32 window, f32 FFT, f32 projection: the
    reference's arithmetic) must sit at that distance or closer to the torch restatement."""
    guard = 2.0 ** -24
    want = nemo_filterbank_features(jfk, 128, 0.97, guard, normalize="per_feature")
    got = o.batch_log_mel(jfk, n_mels=128, preemphasis=0.97, log_zero_guard=guard, normalize_per_feature=True, dtype=np.float32)
    d = np.abs(got.astype(np.float64) - want)
    corr = np.corrcoef(got.reshape(-1), want.reshape(-1))[0, 1]
    assert d.mean() <= 0.0012 and corr >= 0.9997, (d.mean(), corr)
