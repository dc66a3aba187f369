#!/usr/bin/env python
"""bench.py — mel frames/sec of the fused log-mel hot path on N B200s, next to the CPU restatement of the reference.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg4shard|cfg3]

One "step" = one pass of the hot path over one batch of synthetic 16 kHz PCM (BASELINE.json configs[1]:
1024 clips x 10 s, Whisper 80-mel, fft 400 / hop 160).  With N > 1 (launched under torchrun, one rank per GPU)
every rank processes its own shard of that size — clips are independent, there is no data-path collective —
and `value` = frames of all ranks / max-over-ranks device time ("scaling": "weak").

  value      device-resident throughput (CUDA events on the launch stream, inputs already in HBM, 655 MB > L2)
  e2e        same metric through the host-buffer C-ABI call (pinned host PCM -> H2D -> kernel -> D2H) per step
  roofline   algorithmic bytes per launch (4*S + 4*80*F per clip) / measured launch time vs
             MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  oracle/ C restatement of src/stft.rs + src/mel.rs on the box's host cores, bounded sample

`--impl reference` times that same CPU restatement (the reference's Rust cannot be built in this image: no
cargo/rustc) with all host threads on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (clips per GPU, samples per clip, frontend, description)
    "cfg2": (1024, 160000, "whisper", "BASELINE configs[1]: 1024 x 10 s @16 kHz, Whisper 80-mel fft400 hop160"),
    "cfg4shard": (1024, 480000, "whisper", "BASELINE configs[3] per-GPU shard: 1024 x 30 s @16 kHz, Whisper 80-mel"),
    "cfg3": (1024, 160000, "kaldi", "BASELINE configs[2]: 1024 x 10 s @16 kHz, Kaldi 80-bin fbank + CMN"),
}


def measured_traffic(workload):
    """DRAM bytes per step of the hot path from the committed ncu capture (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f).get(workload)
        return (t["traffic"], t["kernel"], t["source"]) if t else (None, None, None)
    except Exception:
        return None, None, None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"      # /opt/skills/guides/B200_PROFILING.md fallback


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def synth_batch_torch(torch, n_clips, n_samples, device, rank):
    """Device-side synthetic PCM of SURVEY §8d's recipe (4 detuned tones + noise, every 8th clip silent for 1 s)."""
    g = torch.Generator(device=device)
    g.manual_seed(1234 + rank)
    t = torch.arange(n_samples, device=device, dtype=torch.float32) / 16000.0
    x = torch.empty((n_clips, n_samples), device=device, dtype=torch.float32)
    blk = 128
    for c0 in range(0, n_clips, blk):
        nb = min(blk, n_clips - c0)
        acc = 0.01 * torch.randn((nb, n_samples), device=device, generator=g)
        for f0, amp in ((220.0, 0.6), (440.0, 0.25), (880.0, 0.10), (1760.0, 0.05)):
            det = 1.0 + (torch.rand((nb, 1), device=device, generator=g) - 0.5) * 0.1
            ph = torch.rand((nb, 1), device=device, generator=g) * 6.283185307
            acc += amp * torch.sin(6.283185307 * f0 * det * t[None, :] + ph)
        x[c0:c0 + nb] = acc
    x[7::8, :16000] = 0.0
    return x


def run_reference(args, rank):
    """CPU arm: the oracle's C restatement, all host threads, bounded sample of the same workload."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import melspec_oracle as o
    import oracle_c as oc
    clips, n_samples, frontend, desc = WORKLOADS[args.workload]
    cores = len(os.sched_getaffinity(0)) or 1   # the host threads this process may actually use
    sample_clips = max(8 * cores, 128)          # ~0.3-1 s of host work per step with all cores
    pcm = np.stack([o.synth_clip(i, n_samples) for i in range(min(sample_clips, 16))])
    pcm = np.ascontiguousarray(np.tile(pcm, (sample_clips // pcm.shape[0] + 1, 1))[:sample_clips])
    fn = (lambda: oc.whisper_batch(pcm, threads=cores)) if frontend == "whisper" else (lambda: oc.kaldi_batch(pcm, threads=cores))
    for _ in range(max(args.warmup, 1)):
        out = fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = fn()
    dt = (time.perf_counter() - t0) / args.steps
    frames = out.shape[0] * out.shape[1]
    v = frames / dt
    sample = f"{sample_clips} clips x {n_samples / 16000:.0f} s per step ({frames} frames), {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": "mel frames/sec (Whisper 80-mel, 16 kHz)", "value": v, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "note": "reference Rust cannot be built here (no cargo); C f64 restatement of "
                   "src/stft.rs+src/mel.rs from oracle/ timed on host cores"},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def bind_to_gpu_cpus(index: int):
    """N > 1: pin this rank to the CPUs NVML reports as local to its GPU before any pinned host memory is allocated, so that
    the end-to-end leg's staging buffers are first-touched on the GPU's own NUMA node (every rank otherwise lands wherever
    the scheduler puts it and the H2D / D2H copies of several ranks share one socket's memory and inter-socket links).
    Not used at N = 1, where the CPU baseline needs all host cores.  Returns the CPU count bound to, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def cpu_baseline(workload):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import melspec_oracle as o
    import oracle_c as oc
    clips, n_samples, frontend, _ = WORKLOADS[workload]
    cores = len(os.sched_getaffinity(0)) or 1   # the host threads this process may actually use
    nclips = max(cores * 2, 16)
    base = np.stack([o.synth_clip(i, n_samples) for i in range(8)])
    pcm = np.ascontiguousarray(np.tile(base, (nclips // 8 + 1, 1))[:nclips])
    fn = (lambda th: oc.whisper_batch(pcm, threads=th)) if frontend == "whisper" else (lambda th: oc.kaldi_batch(pcm, threads=th))
    fn(cores)
    reps, t0 = 0, time.perf_counter()
    while True:
        out = fn(cores)
        reps += 1
        dt = time.perf_counter() - t0
        if dt > 6.0 or reps >= 50:
            break
    v_all = reps * out.shape[0] * out.shape[1] / dt
    one = pcm[:max(2, nclips // cores)]
    t0 = time.perf_counter()
    o1 = oc.whisper_batch(one, threads=1) if frontend == "whisper" else oc.kaldi_batch(one, threads=1)
    v_one = o1.shape[0] * o1.shape[1] / (time.perf_counter() - t0)
    return {"value": v_all, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{nclips} clips x {n_samples / 16000:.0f} s x {reps} reps on {cores} threads "
                      f"(single thread: {v_one:.0f} frames/s)", "single_thread_value": v_one}


def copy_ceiling_ms(torch, hx, hout, dx, dout, reps=3):
    """Copy-only ceiling of the end-to-end call: the same H2D and D2H bytes, pinned, on two streams at once, nothing else."""
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    best = None
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            dx.copy_(hx, non_blocking=True)
        with torch.cuda.stream(s2):
            hout.copy_(dout, non_blocking=True)
        s1.synchronize(); s2.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        best = dt if best is None else min(best, dt)
    return best


def measure_workload(torch, dist, ms, name, args, rank, local_rank, world, steps, sample_clocks):
    """Device-resident timing (CUDA events on the launch stream, max over ranks) + the end-to-end host call of one workload."""
    dev = torch.device("cuda", local_rank)
    clips, n_samples, frontend, desc = WORKLOADS[name]
    if frontend == "whisper":
        h = ms.CudaMelSpectrogram(400, 160, 16000.0, 80, device=local_rank)
    else:
        h = ms.Fbank(ms.FbankConfig(), device=local_rank)
    n_mels = 80
    F = h.num_frames(n_samples)
    x = synth_batch_torch(torch, clips, n_samples, dev, rank)
    out = torch.empty((clips, F, n_mels), dtype=torch.float32, device=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()

    def step():
        h.compute_device(x, clips, n_samples, n_samples, out, stream=stream)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0 and sample_clocks:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = h.launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    with torch.cuda.stream(stream):
        evs[0].record(stream)
        for i in range(steps):
            step()
            evs[i + 1].record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = h.launch_count() - l0
    total_ms = evs[0].elapsed_time(evs[-1])
    per = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(steps))
    tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms_max = float(tt.item())
    # The timed region lasts a few milliseconds, shorter than nvidia-smi's sampling period: keep issuing the very same
    # launches (untimed) until the sampler has seen ~1.5 s of this load, so that the clocks line describes the kernel
    # under load rather than an idle GPU.  Nothing of this continuation enters `value`.
    clocks = None
    if rank == 0 and sample_clocks:
        t_load = time.perf_counter()
        while time.perf_counter() - t_load < 1.5:
            with torch.cuda.stream(stream):
                for _ in range(200 if n_samples < 300000 else 60):
                    step()
            stream.synchronize()
        clocks = sampler.stop()
        clocks["sampled"] = "timed region + 1.5 s untimed continuation of the same launches (nvidia-smi -lms 100)"
    if world > 1:
        dist.barrier()

    # ---- optional: NCCL all-gather of the output shards (north_star: not on the hot path; reported on its own)
    gather = None
    if args.gather and world > 1:
        from mel_spec_b200.shard import gather_output
        full = gather_output(out, world * clips, dist)   # warm-up (NCCL channel setup)
        torch.cuda.synchronize()
        dist.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(3):
            full = gather_output(out, world * clips, dist)
        g1.record()
        torch.cuda.synchronize()
        tg = torch.tensor([g0.elapsed_time(g1) / 3], dtype=torch.float64, device=dev)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        ok = bool(torch.equal(full[rank * clips:(rank + 1) * clips], out))
        recv = (world - 1) * out.numel() * 4
        gather = {"ms": float(tg.item()), "bytes_received_per_gpu": recv, "GB/s_per_gpu": recv / (float(tg.item()) * 1e-3) / 1e9,
                  "own_shard_intact": ok, "collective": "ncclAllGather via torch.distributed.all_gather_into_tensor"}
        del full

    # ---- end to end through the host-buffer C-ABI call (pinned host memory, H2D + kernel + D2H inside the region)
    e2e_steps = max(1, args.e2e_steps)
    hx = torch.empty((clips, n_samples), dtype=torch.float32, pin_memory=True)
    hx.copy_(x)
    hout = torch.empty((clips, F, n_mels), dtype=torch.float32, pin_memory=True)
    torch.cuda.synchronize()
    h.compute_host_raw(hx.data_ptr(), clips, n_samples, n_samples, hout.data_ptr())   # warm-up (allocates staging)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        h.compute_host_raw(hx.data_ptr(), clips, n_samples, n_samples, hout.data_ptr())
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    same = bool(torch.equal(hout.to(dev), out))
    # copy-only ceiling of that call, all ranks at once (the same pinned buffers, H2D and D2H concurrently)
    if world > 1:
        dist.barrier()
    scratch = torch.empty_like(x)
    ceil_ms = copy_ceiling_ms(torch, hx, hout, scratch, out)
    tc = torch.tensor([ceil_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
    ceil_ms = float(tc.item())
    del scratch
    # opt-in 16-bit PCM entry: half the H2D bytes, converted in the kernel's prologue (bit-identical to f32 input of the same samples)
    e2e_i16 = None
    e2e_tga = None
    if frontend == "whisper" and hasattr(ms.lib(), "melspec_compute_host_i16"):
        hx16 = torch.empty((clips, n_samples), dtype=torch.int16, pin_memory=True)
        hx16.copy_((x * 32767.0).round().clamp_(-32768, 32767).to(torch.int16))
        hout16 = torch.empty((clips, F, n_mels), dtype=torch.float32, pin_memory=True)
        h.compute_host_i16_raw(hx16.data_ptr(), clips, n_samples, n_samples, hout16.data_ptr())
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            h.compute_host_i16_raw(hx16.data_ptr(), clips, n_samples, n_samples, hout16.data_ptr())
        t16 = (time.perf_counter() - t0) / e2e_steps
        t16t = torch.tensor([t16], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t16t, op=dist.ReduceOp.MAX)
        t16 = float(t16t.item())
        # identical to the f32 path fed the same (int16 / 32768) samples
        xf = (hx16.to(dev).to(torch.float32) / 32768.0).contiguous()
        ref16 = torch.empty_like(out)
        torch.cuda.synchronize()       # xf was produced on torch's default stream, the launch below is on a non-blocking stream
        h.compute_device(xf, clips, n_samples, n_samples, ref16, stream=stream)
        torch.cuda.synchronize()
        e2e_i16 = {"value": world * clips * F / t16, "unit": "frames/s", "ms_per_step": t16 * 1e3,
                   "h2d_bytes_per_step": clips * n_samples * 2, "d2h_bytes_per_step": clips * F * n_mels * 4,
                   "matches_f32_path_bit_exact": bool(torch.equal(hout16.to(dev), ref16)),
                   "entry": "melspec_compute_host_i16 (opt-in: int16 PCM, x/32768 in the kernel prologue)"}
        # ... and the 8-bit TGA output variant (whisper.cpp's image format, src/quant.rs:38-64): int16 PCM in, one byte per mel value out
        if hasattr(ms.lib(), "melspec_mel_tga_host_batch_i16"):
            width = h.interleaved_width(n_samples)
            tsz = int(ms.lib().melspec_tga_size(n_mels, width))
            htga = torch.empty((clips, tsz), dtype=torch.uint8, pin_memory=True)
            h.mel_tga_batch_raw(hx16.data_ptr(), clips, n_samples, n_samples, htga.data_ptr(), int16=True)
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                h.mel_tga_batch_raw(hx16.data_ptr(), clips, n_samples, n_samples, htga.data_ptr(), int16=True)
            tt = (time.perf_counter() - t0) / e2e_steps
            ttt = torch.tensor([tt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ttt, op=dist.ReduceOp.MAX)
            tt = float(ttt.item())
            # the same bytes as quantising the device path's own frames of clip 0 (interleaved layout) with the device quantiser
            img0 = torch.empty((n_mels, width), dtype=torch.float32, device=dev)
            tga0 = torch.empty((tsz,), dtype=torch.uint8, device=dev)
            torch.cuda.synchronize()
            h.compute_interleaved_device(xf, 1, n_samples, n_samples, 0, img0, stream=stream)
            h.quantize_tga_device(img0, 1, n_mels, width, tga0, stream=stream)
            torch.cuda.synchronize()
            e2e_tga = {"value": world * clips * F / tt, "unit": "frames/s", "ms_per_step": tt * 1e3,
                       "h2d_bytes_per_step": clips * n_samples * 2, "d2h_bytes_per_step": clips * tsz,
                       "clip0_equals_device_quantiser": bool(torch.equal(htga[0].to(dev), tga0)),
                       "entry": "melspec_mel_tga_host_batch_i16 (opt-in: int16 PCM in, 8-bit TGA image per clip out)"}
            del htga, img0, tga0
        del hx16, hout16, xf, ref16
        h.compute_device(x, clips, n_samples, n_samples, out, stream=stream)
        torch.cuda.synchronize()

    frames_rank = clips * F
    ms_per_step = total_ms_max / steps
    algo_bytes = clips * (4 * n_samples + 4 * n_mels * F)
    kern_ms = total_ms / steps                       # this rank's own average launch duration
    peak, peak_kind = measured_peak_gbs()
    achieved = algo_bytes / (kern_ms * 1e-3) / 1e9
    traffic, kname, tsrc = measured_traffic(name)
    res = {
        "metric": "mel frames/sec (Whisper 80-mel, 16 kHz)" if frontend == "whisper" else "fbank frames/sec (Kaldi 80-bin, 16 kHz)",
        "value": world * frames_rank / (ms_per_step * 1e-3), "unit": "frames/s", "ms_per_step": ms_per_step,
        "config": {"workload": desc, "clips_per_gpu": clips, "samples_per_clip": n_samples, "frames_per_clip": F,
                   "l2_policy": "inputs larger than L2 (%.0f MB PCM per launch, no flush)" % (clips * n_samples * 4 / 1e6),
                   "ms_per_step_median_rank0": per[len(per) // 2], "ms_per_step_min_rank0": per[0]},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": tsrc, "kernel": kname, "peak_source": peak_kind,
                     "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": kern_ms,
                     "note": "kernel is shared-memory-wavefront / fp32-issue bound, not HBM bound (DESIGN.md section 3)"},
        "e2e": {"value": world * frames_rank / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": clips * n_samples * 4,
                "d2h_bytes_per_step": clips * F * n_mels * 4, "ms_per_step": e2e_s * 1e3, "matches_device_path": same,
                "copy_ceiling_ms": ceil_ms, "frac_of_copy_ceiling": ceil_ms / (e2e_s * 1e3),
                "copy_ceiling_note": "same pinned buffers, H2D + D2H concurrently on two streams, all ranks at once, no kernel"},
        "gpu_launches": int(launches), "steps": steps,
    }
    if e2e_i16 is not None:
        res["e2e_int16_pcm"] = e2e_i16
        if e2e_tga is not None:
            res["e2e_int16_pcm_tga_out"] = e2e_tga
    if clocks is not None:
        res["clocks"] = clocks
    if gather is not None:
        res["gather"] = gather
    h.close()
    del x, out, hx, hout
    torch.cuda.empty_cache()
    return res


def measure_stream(torch, ms):
    """BASELINE configs[4]: one 1-hour 16 kHz stream.  Device-resident: the whole hour as one clip through the fused kernel
    (roofline).  End to end: the streaming C ABI (overlap-and-save, chunked H2D on a side stream) at three push sizes, and the
    reference's own call shape (one blocking compute_host call), all from pinned host buffers."""
    import ctypes as C
    dev = torch.device("cuda", 0)
    n = 16000 * 3600
    g = torch.Generator().manual_seed(5)
    x = (0.1 * torch.randn(n, generator=g)).pin_memory()
    h = ms.CudaMelSpectrogram(400, 160, 16000.0, 80)
    L = ms.lib()
    F = h.num_frames(n)
    dx = x.to(dev)
    dout = torch.empty((F, 80), dtype=torch.float32, device=dev)
    st = torch.cuda.Stream(device=dev)
    for _ in range(3):
        h.compute_device(dx, 1, n, n, dout, stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 20
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(steps):
            h.compute_device(dx, 1, n, n, dout, stream=st)
        e1.record(st)
    torch.cuda.synchronize()
    kms = e0.elapsed_time(e1) / steps
    peak, peak_kind = measured_peak_gbs()
    algo = 4 * n + 4 * 80 * F
    res = {"config": {"workload": "BASELINE configs[4]: one 1 h @16 kHz stream (57.6 M samples), Whisper 80-mel fft400 hop160",
                      "l2_policy": "input larger than L2 (230 MB PCM per launch, no flush)"},
           "value": F / (kms * 1e-3), "unit": "frames/s", "ms_per_step": kms,
           "roofline": {"bound": "hbm", "achieved": algo / (kms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": algo / (kms * 1e-3) / 1e9 / peak, "peak_source": peak_kind, "algorithmic_bytes_per_launch": algo,
                        "traffic": None, "note": "whole hour device-resident as one clip, one launch"}}
    frames = n // 160 - 3 + 1      # stream framing: first frame at sample 80, whole hops only (src/rb.rs:86-121)
    out = torch.empty((frames + 8, 80), dtype=torch.float32).pin_memory()
    e2e = {}
    ref = None
    for chunk in (16000, 960000, n):
        s = C.c_void_p()
        assert L.melspec_stream_create(h._h, chunk, C.byref(s)) == 0, ms.last_error()
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
        if ref is None:
            ref = out[:frames].clone()
            same = True
        else:
            same = bool(torch.equal(out[:frames], ref))
        e2e[f"push_{chunk // 16000}_s"] = {"value": got / dt, "unit": "frames/s", "seconds": dt, "x_realtime": 3600.0 / dt,
                                          "pushes": (n + chunk - 1) // chunk, "identical_to_1_s_pushes": same,
                                          "h2d_GB/s": 4 * n / dt / 1e9}
        L.melspec_stream_destroy(s)
    # stream frames == batch frames of x[80:]: the device-resident launch above on the shifted signal
    dshift = dx[80:].contiguous()
    h.compute_device(dshift, 1, n - 80, n - 80, dout, stream=st)
    torch.cuda.synchronize()
    e2e["stream_matches_device_path"] = bool(torch.equal(ref.to(dev), dout[:frames]))
    e2e["stream_vs_device_path_max_abs"] = float((ref.to(dev) - dout[:frames]).abs().max().item())
    outb = torch.empty((F, 80), dtype=torch.float32).pin_memory()
    for rep in range(3):
        t0 = time.perf_counter()
        h.compute_host_raw(x.data_ptr(), 1, n, n, outb.data_ptr())
        dt = time.perf_counter() - t0
    h.compute_device(dx, 1, n, n, dout, stream=st)
    torch.cuda.synchronize()
    e2e["one_compute_host_call"] = {"value": F / dt, "unit": "frames/s", "seconds": dt, "x_realtime": 3600.0 / dt,
                                    "h2d_GB/s": 4 * n / dt / 1e9, "matches_device_path": bool(torch.equal(outb.to(dev), dout))}
    res["e2e"] = e2e
    h.close()
    return res


def measure_next_rows(torch, ms, local_rank, steps=10):
    """Device-resident timing of the frontends and sizes beside the headline configuration (SURVEY 8f rows and the general plan),
    same clips as cfg2 (1024 x 10 s), CUDA events on the launch stream: ms per launch, frames/s, fraction of the HBM roofline."""
    dev = torch.device("cuda", local_rank)
    clips, n = 1024, 160000
    x = synth_batch_torch(torch, clips, n, dev, 0)
    st = torch.cuda.Stream(device=dev)
    peak, kind = measured_peak_gbs()
    rows = [
        ("plan 512: Whisper fft 512 hop 160 80 mel (golden-file configuration)", lambda: ms.CudaMelSpectrogram(512, 160, 16000.0, 80, device=local_rank), 80),
        ("plan 400: Whisper large-v3 style, fft 400 hop 160 128 mel", lambda: ms.CudaMelSpectrogram(400, 160, 16000.0, 128, device=local_rank), 128),
        ("plan 512: NeMo BatchLogMel 128 mel, pre-emphasis, per-feature normalisation",
         lambda: ms.BatchLogMelSpectrogram(ms.BatchLogMelConfig(n_mels=128, preemphasis=0.97, normalize_per_feature=True), device=local_rank), 128),
        ("general plan: Whisper fft 1024 hop 256 128 mel", lambda: ms.CudaMelSpectrogram(1024, 256, 16000.0, 128, device=local_rank), 128),
        ("general plan: Whisper fft 480 hop 160 80 mel", lambda: ms.CudaMelSpectrogram(480, 160, 16000.0, 80, device=local_rank), 80),
        ("general plan: Kaldi fbank 8 kHz (fft 256, shift 80) 40 bins + CMN", lambda: ms.Fbank(ms.FbankConfig(sample_rate=8000.0, num_mel_bins=40), device=local_rank), 40),
    ]
    out_rows = []
    for name, mk, nm in rows:
        h = mk()
        F = h.num_frames(n)
        nemo = hasattr(h, "padded_frames")
        cols = h.padded_frames(n) if nemo else F
        o = torch.empty((clips, nm, cols) if nemo else (clips, F, nm), dtype=torch.float32, device=dev)
        lay = ms.LAYOUT_MEL_MAJOR if nemo else ms.LAYOUT_FRAME_MAJOR
        for _ in range(3):
            h.compute_device(x, clips, n, n, o, layout=lay, stream=st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record(st)
            for _ in range(steps):
                h.compute_device(x, clips, n, n, o, layout=lay, stream=st)
            e1.record(st)
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / steps
        algo = clips * (4 * n + 4 * nm * cols)
        out_rows.append({"row": name + ", 1024 x 10 s", "ms_per_step": t, "value": clips * F / (t * 1e-3), "unit": "frames/s",
                         "frames_per_clip": F, "roofline_frac": algo / (t * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": algo,
                         "peak_source": kind})
        h.close()
        del o
    return out_rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: cfg2 (BASELINE configs[1]) on one GPU, cfg4shard (configs[3], 1024 x 30 s per GPU) under torchrun")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="N = 1: skip the `extra` block (the other BASELINE configs)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--gather", action="store_true",
                    help="N > 1: also time the optional NCCL all-gather of the output shards (reported separately, never in `value`)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    explicit = args.workload is not None
    if args.workload is None:
        args.workload = "cfg2" if world == 1 else "cfg4shard"

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import mel_spec_b200 as ms

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback")
    numa = bind_to_gpu_cpus(local_rank) if world > 1 else None
    ms.build()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    res = measure_workload(torch, dist, ms, args.workload, args, rank, local_rank, world, args.steps, True)
    if rank == 0:
        line = {
            "metric": res["metric"], "value": res["value"], "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": res["config"], "roofline": res["roofline"],
            "e2e": res["e2e"], "gpu_launches": res["gpu_launches"], "clocks": res.get("clocks"),
        }
        for k in ("e2e_int16_pcm", "e2e_int16_pcm_tga_out", "gather"):
            if k in res:
                line[k] = res[k]
        if numa is not None:
            line["config"]["host_affinity"] = f"each rank bound to the {numa} CPUs local to its GPU (NVML) before pinned allocation"
        if world == 1 and not args.no_extra and not explicit:
            # the other BASELINE configs on this GPU, same method (fewer steps): configs[2] Kaldi, configs[3]'s per-GPU shard,
            # configs[4] the 1-hour stream.  configs[0] (JFK vs golden) is a parity case: tests/ and smoke().
            extra = {}
            for wl in ("cfg3", "cfg4shard"):
                r = measure_workload(torch, dist, ms, wl, args, 0, local_rank, 1, max(5, args.steps // 2), False)
                extra[wl] = {k: r[k] for k in ("metric", "value", "unit", "ms_per_step", "config", "roofline", "e2e", "gpu_launches", "steps")
                             if k in r}
                for k in ("e2e_int16_pcm", "e2e_int16_pcm_tga_out"):
                    if k in r:
                        extra[wl][k] = r[k]
            extra["cfg5_stream"] = measure_stream(torch, ms)
            extra["next_rows"] = measure_next_rows(torch, ms, local_rank)
            line["extra"] = extra
        if not args.no_cpu_baseline and world == 1:      # reported baseline: rank 0 at N = 1 only
            line["cpu_baseline"] = cpu_baseline(args.workload)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
