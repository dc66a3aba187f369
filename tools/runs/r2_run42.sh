#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O; rm -f $O/run42.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/run42.txt
cat > /tmp/w512.py <<'P'
import sys, os, torch
sys.path.insert(0, ".")
import mel_spec_b200 as ms
from bench import synth_batch_torch
from tools.bench_next_rows import timeit
dev = torch.device("cuda", 0)
clips, n = 1024, 160000
x = synth_batch_torch(torch, clips, n, dev, 0)
st = torch.cuda.Stream(device=dev)
h = ms.CudaMelSpectrogram(512, 160, 16000.0, 80)
F = h.num_frames(n)
res = []
for lay in (0, 1):
    o = torch.empty((clips, F, 80) if lay == 0 else (clips, 80, F), dtype=torch.float32, device=dev)
    t = timeit(lambda: h.compute_device(x, clips, n, n, o, layout=lay, stream=st), 20, st)
    res.append(f"layout {lay}: {t:.4f} ms")
print("KSCHED=" + os.environ.get("MELSPEC_KSCHED", "default"), "whisper512", " | ".join(res))
P
for i in 1 2; do MELSPEC_KSCHED=0 timeout 300 python /tmp/w512.py >> $O/run42.txt 2>&1; timeout 300 python /tmp/w512.py >> $O/run42.txt 2>&1; done
cat $O/run42.txt
