#!/usr/bin/env python
"""Two launches each of the NeMo frontend (80 / 128 mel) and Whisper-512 mel-major at 1024 x 10 s, for an ncu metrics pass."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mel_spec_b200 as ms
from bench import synth_batch_torch
dev = torch.device("cuda", 0)
clips, n = 1024, 160000
x = synth_batch_torch(torch, clips, n, dev, 0)
for mk, nm in ((lambda: ms.BatchLogMelSpectrogram(ms.BatchLogMelConfig(n_mels=80)), 80),
               (lambda: ms.BatchLogMelSpectrogram(ms.BatchLogMelConfig(n_mels=128, preemphasis=0.97)), 128),
               (lambda: ms.CudaMelSpectrogram(512, 160, 16000.0, 80), 80)):
    h = mk()
    F = h.num_frames(n)
    cols = h.padded_frames(n) if hasattr(h, "padded_frames") else F
    o = torch.empty((clips, nm, cols), dtype=torch.float32, device=dev)
    for _ in range(2):
        h.compute_device(x, clips, n, n, o, layout=1)
    torch.cuda.synchronize()
    del o
    h.close()
