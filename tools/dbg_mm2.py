#!/usr/bin/env python
"""Debug: run-to-run determinism of one layout / mel count over many launches (compared on the device)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import mel_spec_b200 as ms
import melspec_oracle as o
n_mels, lay, reps, nfr = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
h = ms.CudaMelSpectrogram(400, 160, 16000.0, n_mels)
clips, n = 300, 400 + 160 * (nfr - 1)
pcm = np.stack([o.synth_clip(i % 7, n) * (0.1 + 0.05 * (i % 5)) for i in range(clips)]).astype(np.float32)
x = torch.from_numpy(pcm).cuda()
f = h.num_frames(n)
shape = (clips, f, n_mels) if lay == 0 else (clips, n_mels, f)
ref = torch.empty(shape, dtype=torch.float32, device="cuda")
h.compute_device(x, clips, n, n, ref, layout=lay)
torch.cuda.synchronize()
bad = 0
where = []
for rep in range(reps):
    out = torch.full(shape, float("nan"), dtype=torch.float32, device="cuda")
    h.compute_device(x, clips, n, n, out, layout=lay)
    torch.cuda.synchronize()
    ne = (out != ref)
    if bool(ne.any()):
        bad += 1
        idx = ne.nonzero()[:1].cpu().numpy()[0]
        where.append(tuple(int(v) for v in idx))
        if bad <= 3 and lay == 1:
            c, fr = int(idx[0]), int(idx[2])
            fr0 = fr - (fr % 6)
            d = (out[c, :, fr0:fr0 + 6] - ref[c, :, fr0:fr0 + 6]).cpu().numpy()
            np.set_printoptions(linewidth=250, precision=2)
            print("clip", c, "tile frames", fr0, ".. +5; per-frame count of differing mels:", (d != 0).sum(axis=0), " max |d| per frame:", np.abs(d).max(axis=0))
            q = int(np.argmax(np.abs(d).max(axis=0)))
            print("   d[mel 0..23, frame", fr0 + q, "] =", d[:24, q])
            print("   ref[mel 0..11] =", ref[c, :12, fr0 + q].cpu().numpy(), " mx-related: ref min/max over mels", float(ref[c, :, fr0 + q].min()), float(ref[c, :, fr0 + q].max()))
print(f"n_mels {n_mels} layout {lay} frames {f}: {bad} of {reps} launches differ from the first", where[:6],
      {k: os.environ[k] for k in os.environ if k.startswith("MELSPEC_")})
