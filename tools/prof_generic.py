#!/usr/bin/env python
"""One general-plan configuration, a few launches (for ncu): python tools/prof_generic.py fft hop n_mels"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mel_spec_b200 as ms
from bench import synth_batch_torch
fft, hop, nm = (int(a) for a in sys.argv[1:4])
dev = torch.device("cuda", 0)
clips, n = 1024, 160000
x = synth_batch_torch(torch, clips, n, dev, 0)
h = ms.CudaMelSpectrogram(fft, hop, 16000.0, nm)
F = h.num_frames(n)
o = torch.empty((clips, F, nm), dtype=torch.float32, device=dev)
for _ in range(3):
    h.compute_device(x, clips, n, n, o)
torch.cuda.synchronize()
print("frames", clips * F)
