import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, numpy as np
import mel_spec_b200 as ms
from bench import synth_batch_torch
dev = torch.device("cuda", 0)
h = ms.CudaMelSpectrogram(400, 160, 16000.0, 80)
for clips in (3, 60, 200, 1024):
    n = 160000
    x = synth_batch_torch(torch, clips, n, dev, 0)
    x16 = (x * 32767.0).round().clamp_(-32768, 32767).to(torch.int16)
    hx16 = torch.empty((clips, n), dtype=torch.int16, pin_memory=True); hx16.copy_(x16)
    F = h.num_frames(n)
    hout = torch.empty((clips, F, 80), dtype=torch.float32, pin_memory=True)
    h.compute_host_i16_raw(hx16.data_ptr(), clips, n, n, hout.data_ptr())
    xf = (x16.to(torch.float32) / 32768.0).contiguous()
    ref = torch.empty((clips, F, 80), dtype=torch.float32, device=dev)
    h.compute_device(xf, clips, n, n, ref)
    torch.cuda.synchronize()
    hf = torch.empty((clips, F, 80), dtype=torch.float32, pin_memory=True)
    hxf = xf.cpu().pin_memory()
    h.compute_host_raw(hxf.data_ptr(), clips, n, n, hf.data_ptr())
    d = (hout.to(dev) - ref).abs()
    d2 = (hf.to(dev) - ref).abs()
    bad = (d > 0).nonzero()
    print(clips, "i16 vs dev: max", float(d.max()), "n bad", int((d > 0).sum()), "first", bad[:3].tolist(), "| f32 host vs dev max", float(d2.max()))
