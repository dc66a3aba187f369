"""Where does a small streaming push spend its time?  Times melspec_stream_push for several chunk sizes (pinned and pageable
host buffers) and, for comparison, the bare CUDA operations a push is made of (H2D + empty kernel-sized gap + D2H + sync)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import mel_spec_b200 as ms
h = ms.CudaMelSpectrogram(400, 160, 16000.0, 80)
L = ms.lib()
for pinned in (True, False):
    for chunk in (1600, 16000, 160000):
        x = torch.randn(chunk * 64) * 0.1
        out = torch.empty((chunk // 160 + 8, 80))
        if pinned:
            x, out = x.pin_memory(), out.pin_memory()
        s = C.c_void_p(); assert L.melspec_stream_create(h._h, chunk, C.byref(s)) == 0
        em = C.c_int64(0)
        for rep in range(2):
            L.melspec_stream_reset(s)
            t0 = time.perf_counter()
            for i in range(64):
                rc = L.melspec_stream_push(s, x.data_ptr() + 4 * chunk * i, chunk, out.data_ptr(), out.shape[0], C.byref(em))
                assert rc == 0
            dt = (time.perf_counter() - t0) / 64
        print(f"pinned={pinned} chunk={chunk:7d} samples: {dt*1e6:8.1f} us per push  ({chunk/160/dt/1e6:.2f} M frames/s)")
        L.melspec_stream_destroy(s)
# bare CUDA round trip of the same sizes
st = torch.cuda.Stream()
for chunk in (1600, 16000, 160000):
    x = (torch.randn(chunk) * 0.1).pin_memory(); d = torch.empty(chunk, device="cuda")
    o = torch.empty(chunk // 2, device="cuda"); ho = torch.empty(chunk // 2).pin_memory()
    for rep in range(2):
        t0 = time.perf_counter()
        for i in range(64):
            with torch.cuda.stream(st):
                d.copy_(x, non_blocking=True); o.add_(1.0); ho.copy_(o, non_blocking=True)
            st.synchronize()
        dt = (time.perf_counter() - t0) / 64
    print(f"bare H2D + tiny kernel + D2H + sync, chunk={chunk}: {dt*1e6:.1f} us")
