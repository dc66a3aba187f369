#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O; rm -f $O/run33_ab.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/run33_tests.txt
python - >> $O/run33_ab.txt 2>&1 <<'P'
import os, sys, subprocess
code = r'''
import sys, torch
sys.path.insert(0, ".")
import mel_spec_b200 as ms
from bench import synth_batch_torch
from tools.bench_next_rows import timeit
dev = torch.device("cuda", 0)
clips, n = 1024, 160000
x = synth_batch_torch(torch, clips, n, dev, 0)
st = torch.cuda.Stream(device=dev)
res = []
for nm in (80, 128):
    h = ms.CudaMelSpectrogram(400, 160, 16000.0, nm)
    F = h.num_frames(n)
    for lay in (0, 1):
        o = torch.empty((clips, F, nm) if lay == 0 else (clips, nm, F), dtype=torch.float32, device=dev)
        t = timeit(lambda: h.compute_device(x, clips, n, n, o, layout=lay, stream=st), 20, st)
        res.append(f"{nm} mel layout {lay}: {t:.4f} ms")
        del o
    h.close()
print(" | ".join(res))
'''
for env in ({}, {"MELSPEC_KSPEC5": "0"}, {"MELSPEC_KSPEC5": "0", "MELSPEC_TILE_ORDER": "0"}):
    e = dict(os.environ, **env)
    r = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True)
    print(env, r.stdout.strip(), r.stderr.strip()[-300:])
P
cat $O/run33_tests.txt $O/run33_ab.txt
