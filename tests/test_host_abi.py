"""CPU-only checks of the C-ABI library: it loads, exports every symbol include/melspec_b200.h declares, and its
device-free host logic (filterbanks, framing, config validation, error strings) matches the oracle.  No compute."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import melspec_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def m():
    import mel_spec_b200 as mod
    mod.build()
    return mod


def test_library_exports_every_declared_symbol(m):
    hdr = open(os.path.join(ROOT, "include", "melspec_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(melspec_[a-z_0-9]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    L = m.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert sorted(m.EXPORTS) == declared
    assert L.melspec_abi_version() == 2


def test_filterbanks_match_golden_and_oracle(m, golden_dir):
    assert np.abs(m.mel(16000, 400, 80) - np.load(os.path.join(golden_dir, "mel_filters_80x201.npy"))).max() <= 1e-7
    assert np.abs(m.mel(16000, 512, 80) - np.load(os.path.join(golden_dir, "nemo_filters_80x257.npy"))).max() <= 1e-7
    assert np.abs(m.mel(16000, 400, 80) - o.slaney_mel_filterbank(16000, 400, 80)).max() <= 1e-12
    assert np.abs(m.mel(16000, 512, 128) - o.slaney_mel_filterbank(16000, 512, 128)).max() <= 1e-12
    assert np.abs(m.kaldi_mel_filterbank() - o.kaldi_mel_filterbank()).max() <= 1e-12
    assert m.mel(16000, 400, 80).shape == (80, 201)          # tests/readme_examples.rs:37-38


def test_default_configs_and_frame_counts(m):
    from mel_spec_b200._lib import MelspecConfig
    L = m.lib()
    cfg = MelspecConfig()
    assert L.melspec_default_config(0, C.byref(cfg)) == 0
    assert (cfg.fft_size, cfg.hop_size, cfg.n_mels, cfg.sampling_rate) == (400, 160, 80, 16000.0)
    for n, want in ((0, 0), (399, 0), (400, 1), (16000, 98), (160000, 998), (480000, 2998), (57600000, 359998)):
        assert L.melspec_num_frames_cfg(C.byref(cfg), n) == want == o.num_frames(n, 400, 160)
    assert L.melspec_default_config(1, C.byref(cfg)) == 0
    assert (cfg.frame_length, cfg.hop_size, cfg.preemphasis, cfg.low_freq, cfg.apply_cmn) == (400, 160, 0.97, 20.0, 1)
    assert L.melspec_num_frames_cfg(C.byref(cfg), 176000) == 1098
    assert L.melspec_default_config(7, C.byref(cfg)) != 0


def test_invalid_configs_are_rejected_without_a_device(m):
    from mel_spec_b200._lib import MelspecConfig
    L = m.lib()
    cfg = MelspecConfig()
    L.melspec_default_config(0, C.byref(cfg))
    cfg.n_mels = 0                                           # src/cuda.rs:45-49
    buf = (C.c_double * 8)()
    assert L.melspec_build_filterbank(C.byref(cfg), buf, 8) == 1
    assert b"non-zero" in L.melspec_last_error()
    cfg.n_mels = 80
    assert L.melspec_build_filterbank(C.byref(cfg), buf, 8) == 4      # capacity too small
    with pytest.raises(m.CudaError) as ei:
        m.CudaMelSpectrogram(0, 160, 16000.0, 80)
    assert ei.value.kind == "Unavailable"


def test_no_cpu_fallback(m):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present; the no-device error path is exercised on CPU boxes")
    with pytest.raises(m.CudaError) as ei:
        m.CudaMelSpectrogram(400, 160, 16000.0, 80)
    assert ei.value.kind == "Unavailable" and "no CPU fallback" in str(ei.value)


def test_general_sizes_host_logic(m):
    """Device-free host logic for configurations that run on the general plan: filterbanks and frame counts for arbitrary
    fft_size / hop / sample rate / mel scale match the oracle (src/mel.rs:547-643, src/fbank.rs:253-313, src/stft.rs:153-157)."""
    from mel_spec_b200._lib import MelspecConfig
    from mel_spec_b200 import api
    L = m.lib()
    for sr, fft, n_mels in ((16000, 1024, 128), (8000, 256, 40), (22050, 441, 64), (48000, 8192, 80), (16000, 16, 4), (16000, 251, 40)):
        assert np.abs(m.mel(sr, fft, n_mels) - o.slaney_mel_filterbank(sr, fft, n_mels)).max() <= 1e-12
    for kw in (dict(sample_rate=8000.0, num_mel_bins=40), dict(sample_rate=44100.0, num_mel_bins=64),
               dict(low_freq=100.0, high_freq=7000.0, num_mel_bins=23, sample_rate=22050.0)):
        fc = m.FbankConfig(**kw)
        want = o.kaldi_mel_filterbank(fc.sample_rate, fc.fft_size(), fc.num_mel_bins, fc.low_freq,
                                      fc.high_freq if fc.high_freq else fc.sample_rate / 2)
        assert np.abs(m.kaldi_mel_filterbank(fc) - want).max() <= 1e-12
    # NeMo bank with HTK scale, no area normalisation, band limits
    bc = m.BatchLogMelConfig(n_fft=1024, win_length=800, hop_length=200, n_mels=128, htk=True, norm=False, f_min=50.0, f_max=7600.0)
    cfg = api._nemo_cfg(bc)
    out = np.zeros((128, 513), dtype=np.float64)
    assert L.melspec_build_filterbank(C.byref(cfg), out.ctypes.data_as(C.POINTER(C.c_double)), out.size) == 0
    assert np.abs(out - o.general_mel_filterbank(16000.0, 1024, 128, 50.0, 7600.0, True, False)).max() <= 1e-12
    # frame counts: Whisper (len - N)/hop + 1, NeMo centred len/hop + 1 or (len - n_fft)/hop + 1
    w = api._whisper_cfg(1024, 256, 128, 16000.0)
    for n in (0, 1023, 1024, 1279, 1280, 100000):
        assert L.melspec_num_frames_cfg(C.byref(w), n) == o.num_frames(n, 1024, 256)
    for center in (True, False):
        c2 = api._nemo_cfg(m.BatchLogMelConfig(n_fft=768, win_length=601, hop_length=123, center=center))
        for n in (0, 1, 767, 768, 50000):
            want = 0 if n == 0 else o.batch_num_frames(n, 768, 123, center)
            assert L.melspec_num_frames_cfg(C.byref(c2), n) == want


def test_environment_switches_are_documented():
    """Every MELSPEC_* variable the library reads (std::getenv in csrc/) has a row in INTEGRATION.md's table, and the table
    names no variable the sources do not read (MELSPEC_B200_LIB belongs to the Python loader)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    read = set()
    csrc = os.path.join(root, "mel-spec_b200", "csrc")
    for name in os.listdir(csrc):
        read |= set(re.findall(r'getenv\("(MELSPEC_[A-Z0-9_]+)"\)', open(os.path.join(csrc, name)).read()))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    table = doc[doc.index("## Environment switches of the library"):]
    named = set()
    for cell in re.findall(r"^\| (.+?) \|", table, flags=re.M):
        named |= set(re.findall(r"`(MELSPEC_[A-Z0-9_]+)(?:=\d)?`", cell))
    named.discard("MELSPEC_B200_LIB")
    assert read and read == named, (sorted(read - named), sorted(named - read))
