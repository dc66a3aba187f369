#!/usr/bin/env python
"""Plan-512 modes side by side (device-resident, 1024 x 10 s, CUDA events): Whisper-512, Kaldi without / with CMN, NeMo 80 / 128 mel.
One line per mode: ms per launch.  MELSPEC_B200_LIB selects the build (tools/ab_bench.sh style A/B)."""
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
tag = os.path.basename(os.environ.get("MELSPEC_B200_LIB", "default"))
rows = [("whisper512", lambda: ms.CudaMelSpectrogram(512, 160, 16000.0, 80), 80, 0),
        ("kaldi_nocmn", lambda: ms.Fbank(ms.FbankConfig(apply_cmn=False)), 80, 0),
        ("kaldi_cmn", lambda: ms.Fbank(ms.FbankConfig()), 80, 0),
        ("nemo80", lambda: ms.BatchLogMelSpectrogram(ms.BatchLogMelConfig(n_mels=80)), 80, 1),
        ("nemo128_pre", lambda: ms.BatchLogMelSpectrogram(ms.BatchLogMelConfig(n_mels=128, preemphasis=0.97)), 128, 1),
        ("nemo128_pre_norm", lambda: ms.BatchLogMelSpectrogram(ms.BatchLogMelConfig(n_mels=128, preemphasis=0.97, normalize_per_feature=True)), 128, 1),
        ("whisper400", lambda: ms.CudaMelSpectrogram(400, 160, 16000.0, 80), 80, 0)]
out = []
for name, mk, nm, lay in rows:
    h = mk()
    F = h.num_frames(n)
    cols = h.padded_frames(n) if hasattr(h, "padded_frames") else F
    o = torch.empty((clips, nm, cols) if lay else (clips, F, nm), dtype=torch.float32, device=dev)
    t = timeit(lambda: h.compute_device(x, clips, n, n, o, layout=lay, stream=st), 20, st)
    out.append(f"{name} {t:.4f}")
    h.close()
    del o
print(tag, " | ".join(out))
