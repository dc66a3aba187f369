#!/usr/bin/env python
"""Debug: repeat the frame-major vs mel-major comparison on a ragged batch and report where they differ."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import mel_spec_b200 as ms
import melspec_oracle as o
for n_mels in (128,):
    h = ms.CudaMelSpectrogram(400, 160, 16000.0, n_mels)
    clips, n = 300, 16000 * 2 + 400 + 160 * 3 - 320
    pcm = np.stack([o.synth_clip(i % 7, n) * (0.1 + 0.05 * (i % 5)) for i in range(clips)]).astype(np.float32)
    x = torch.from_numpy(pcm).cuda()
    f = h.num_frames(n)
    ref = None
    for rep in range(60):
        outs = []
        for lay in (0, 1):
            out = torch.full((clips, f, n_mels) if lay == 0 else (clips, n_mels, f), float("nan"), dtype=torch.float32, device="cuda")
            h.compute_device(x, clips, n, n, out, layout=lay)
            torch.cuda.synchronize()
            a = out.cpu().numpy()
            outs.append(a if lay == 0 else a.transpose(0, 2, 1))
        if ref is None:
            ref = outs[0].copy()
        for name, arr in (("frame-major", outs[0]), ("mel-major", outs[1])):
            bad = np.argwhere(~((arr == ref) | (np.isnan(arr) & np.isnan(ref))))
            if bad.size:
                cl = np.unique(bad[:, 0]); fr = np.unique(bad[:, 1]); me = np.unique(bad[:, 2])
                print(n_mels, "rep", rep, name, "mismatches", len(bad), "clips", cl[:8], "frames", fr[:12], "mels", me[:8], "...", me[-4:],
                      "nan in arr", int(np.isnan(arr[tuple(bad.T)]).sum()), "sample", arr[tuple(bad[0])], ref[tuple(bad[0])])
                c0, f0 = int(cl[0]), int(fr[0])
                row = arr[c0, f0]
                src = [(c, f) for c in range(clips) for f in range(ref.shape[1]) if np.array_equal(ref[c, f], row)]
                print("   deviating frame", (c0, f0), "equals ref frames:", src[:10], " max|dev - ref| =", float(np.abs(row - ref[c0, f0]).max()),
                      " amplitude idx of clip:", c0 % 5, "seed idx:", c0 % 7)
    print(n_mels, "done")
    h.close()
