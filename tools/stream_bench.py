#!/usr/bin/env python
"""BASELINE config 5: one 1-hour 16 kHz stream through the streaming C ABI (overlap-and-save, chunked H2D on a side
stream).  Prints frames/s for several push sizes.  Run on a B200: python tools/stream_bench.py"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import mel_spec_b200 as ms

n = 16000 * 3600
g = torch.Generator().manual_seed(5)
x = (0.1 * torch.randn(n, generator=g)).pin_memory()
frames = n // 160 - 3 + 1
out = torch.empty((frames + 8, 80), dtype=torch.float32).pin_memory()
h = ms.CudaMelSpectrogram(400, 160, 16000.0, 80)
L = ms.lib()
res = {}
for chunk in (16000, 160000, 960000, n):
    s = C.c_void_p()
    assert L.melspec_stream_create(h._h, chunk, C.byref(s)) == 0
    for rep in range(2):
        L.melspec_stream_reset(s)
        got, t0 = 0, time.perf_counter()
        for off in range(0, n, chunk):
            m = min(chunk, n - off)
            em = C.c_int64(0)
            rc = L.melspec_stream_push(s, x.data_ptr() + 4 * off, m, out.data_ptr() + 4 * 80 * got, out.shape[0] - got, C.byref(em))
            assert rc == 0, ms.last_error()
            got += em.value
        dt = time.perf_counter() - t0
    assert got == frames, (got, frames)
    res[f"push_{chunk}_samples"] = {"frames_per_s": got / dt, "seconds": dt, "x_realtime": 3600.0 / dt}
    L.melspec_stream_destroy(s)
# the reference's own call shape on the whole hour: one blocking host call (time-split pipeline inside the library)
fb = h.num_frames(n)
outb = torch.empty((fb, 80), dtype=torch.float32).pin_memory()
for rep in range(3):
    t0 = time.perf_counter()
    h.compute_host_raw(x.data_ptr(), 1, n, n, outb.data_ptr())
    dt = time.perf_counter() - t0
res["one_compute_host_call"] = {"frames_per_s": fb / dt, "seconds": dt, "x_realtime": 3600.0 / dt}
print(json.dumps({"workload": "BASELINE configs[4]: 1 h @16 kHz stream, Whisper 80-mel fft400 hop160, pinned host buffers",
                  "frames": frames, "results": res}))
