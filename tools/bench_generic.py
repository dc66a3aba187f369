#!/usr/bin/env python
"""General-plan sizes side by side (device-resident, 1024 x 10 s, CUDA events): one line, ms per launch and M frames/s per size.
MELSPEC_GENERIC_PAIR / MELSPEC_PAIR_MIN_WARPS / MELSPEC_B200_LIB select the kernel form and the build (A/B)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mel_spec_b200 as ms
from bench import synth_batch_torch
from tools.bench_next_rows import timeit
dev = torch.device("cuda", 0)
clips, n = 1024, 160000
x = synth_batch_torch(torch, clips, n, dev, 0)
st = torch.cuda.Stream(device=dev)
rows = [("w1024/256/128", lambda: ms.CudaMelSpectrogram(1024, 256, 16000.0, 128), 128),
        ("w2048/512/80", lambda: ms.CudaMelSpectrogram(2048, 512, 16000.0, 80), 80),
        ("w480/160/80", lambda: ms.CudaMelSpectrogram(480, 160, 16000.0, 80), 80),
        ("w512/128/80", lambda: ms.CudaMelSpectrogram(512, 128, 16000.0, 80), 80),
        ("w400/320/80", lambda: ms.CudaMelSpectrogram(400, 320, 16000.0, 80), 80),
        ("w800/200/80", lambda: ms.CudaMelSpectrogram(800, 200, 16000.0, 80), 80),
        ("w450/150/64", lambda: ms.CudaMelSpectrogram(450, 150, 16000.0, 64), 64),
        ("kaldi8k+cmn", lambda: ms.Fbank(ms.FbankConfig(sample_rate=8000.0, num_mel_bins=40)), 40),
        ("nemo1024/256/80", lambda: ms.BatchLogMelSpectrogram(ms.BatchLogMelConfig(n_fft=1024, win_length=1024, hop_length=256, n_mels=80)), 80)]
out = []
for name, mk, nm in rows:
    h = mk()
    F = h.num_frames(n)
    nemo = hasattr(h, "padded_frames")
    cols = h.padded_frames(n) if nemo else F
    o = torch.empty((clips, nm, cols) if nemo else (clips, F, nm), dtype=torch.float32, device=dev)
    t = timeit(lambda: h.compute_device(x, clips, n, n, o, layout=1 if nemo else 0, stream=st), 10, st)
    out.append(f"{name} {t:.3f} ms {clips * F / t / 1e3:.0f} M")
    h.close()
    del o
tag = "pair=" + os.environ.get("MELSPEC_GENERIC_PAIR", "default") + " minw=" + os.environ.get("MELSPEC_PAIR_MIN_WARPS", "default")
print(tag, " | ".join(out))
