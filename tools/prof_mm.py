#!/usr/bin/env python
"""A few launches of the fused kernel with the mel-major (interleave_frames) layout, for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mel_spec_b200 as ms
from bench import synth_batch_torch
dev = torch.device("cuda", 0)
clips, n = 1024, 160000
x = synth_batch_torch(torch, clips, n, dev, 0)
h = ms.CudaMelSpectrogram(400, 160, 16000.0, 80)
W = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
img = torch.empty((clips, 80, W), dtype=torch.float32, device=dev)
for _ in range(3):
    h.compute_interleaved_device(x, clips, n, n, W if W != 998 else 0, img)
torch.cuda.synchronize()
