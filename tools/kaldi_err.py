import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import numpy as np, mel_spec_b200 as ms, melspec_oracle as o
jfk = np.load("tests/golden/jfk_pcm_f32.npy")
fb = ms.Fbank(ms.FbankConfig()); got = fb.compute(jfk); want = o.kaldi_fbank(jfk); d = np.abs(got - want)
print("kaldi jfk: max %.3e mean %.3e  >1e-4: %.3f%%  >1e-3: %.3f%%" % (d.max(), d.mean(), 100 * (d > 1e-4).mean(), 100 * (d > 1e-3).mean()))
gold = np.load("tests/golden/kaldi_fbank_jfk.npy").T; d = np.abs(got - gold); print("kaldi vs knf golden: max %.3e mean %.3e" % (d.max(), d.mean()))
for fft in (400, 512):
    h = ms.CudaMelSpectrogram(fft, 160, 16000.0, 80); got = h.compute_mel_spectrogram(jfk); d = np.abs(got - o.whisper_mel_batch(jfk, fft)); print("whisper", fft, "jfk max %.3e mean %.3e" % (d.max(), d.mean()))
h = ms.CudaMelSpectrogram(512, 160, 16000.0, 80); g = np.load("tests/golden/rust_jfk_golden.npy"); d = np.abs(h.compute_host(jfk[128:], layout=1) - g); print("whisper512 vs rust golden: max %.3e mean %.3e" % (d.max(), d.mean()))
