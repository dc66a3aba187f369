#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O; rm -f $O/run37.txt
cat > /tmp/lay.py <<'P'
import sys, os, torch
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
    o = torch.empty((clips, nm, F), dtype=torch.float32, device=dev)
    t = timeit(lambda: h.compute_device(x, clips, n, n, o, layout=1, stream=st), 20, st)
    res.append(f"{nm} mel mel-major: {t:.4f} ms")
    del o
    h.close()
print("MM_SYNC=" + os.environ.get("MELSPEC_MM_SYNC", "default"), " | ".join(res))
P
for s in 0 8 12 16 24 32 48; do MELSPEC_MM_SYNC=$s timeout 300 python /tmp/lay.py >> $O/run37.txt 2>&1; done
timeout 300 python /tmp/lay.py >> $O/run37.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_formats.py -m gpu -x -q 2>&1 | tail -3 >> $O/run37.txt
cat $O/run37.txt
