"""Launch the NeMo frontend (128 mel, n_fft 512) on 1024 x 10 s a few times (target for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, mel_spec_b200 as ms
from bench import synth_batch_torch
dev = torch.device("cuda", 0); x = synth_batch_torch(torch, 1024, 160000, dev, 0)
nemo = ms.BatchLogMelSpectrogram(ms.BatchLogMelConfig(n_mels=128)); cols = nemo.padded_frames(160000)
out = torch.empty((1024, 128, cols), dtype=torch.float32, device=dev)
for _ in range(3):
    nemo.compute_device(x, 1024, 160000, 160000, out, layout=1)
torch.cuda.synchronize()
