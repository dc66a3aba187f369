#!/usr/bin/env python
"""One launch per (n_mels, layout) of plan 400 at 1024 x 10 s, for an ncu metrics pass."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mel_spec_b200 as ms
from bench import synth_batch_torch
dev = torch.device("cuda", 0)
clips, n = 1024, 160000
x = synth_batch_torch(torch, clips, n, dev, 0)
for nm in (80, 128):
    h = ms.CudaMelSpectrogram(400, 160, 16000.0, nm)
    F = h.num_frames(n)
    for lay in (0, 1):
        o = torch.empty((clips, F, nm) if lay == 0 else (clips, nm, F), dtype=torch.float32, device=dev)
        for _ in range(2):
            h.compute_device(x, clips, n, n, o, layout=lay)
        torch.cuda.synchronize()
        del o
    h.close()
