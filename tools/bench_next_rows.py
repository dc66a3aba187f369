#!/usr/bin/env python
"""Measurement of the SURVEY §8f rows (the callers / formats either side of the hot path) on one B200, device-resident,
CUDA events on the launch stream, inputs larger than L2.  One JSON object per row on stdout:
  f-1  NeMo BatchLogMel frontend (128 mel, n_fft 512, per-feature normalisation), 1024 clips x 10 s
  f-3a fused kernel writing the interleave_frames image directly (mel-major, min_width padding)
  f-3b TGA quantiser (min/max pass + quantise pass) and dequantiser on 1024 images of 80 x 998
  f-4  VAD Sobel/majority kernel + per-frame activity kernel on the same images
  general plan: Whisper at fft 1024 / 2048 / 480 and Kaldi at 8 kHz (sizes outside the specialised kernels)
`achieved` = algorithmic bytes / time against MEASURED_PEAKS.json hbm_gbs (these are HBM-bound byte/stencil kernels).
Run: python tools/bench_next_rows.py [--steps 20]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mel_spec_b200 as ms
from bench import measured_peak_gbs, synth_batch_torch


def timeit(fn, steps, stream):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for _ in range(steps):
            fn()
        ev[1].record(stream)
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    peak, kind = measured_peak_gbs()
    clips, n = 1024, 160000
    x = synth_batch_torch(torch, clips, n, dev, 0)
    st = torch.cuda.Stream(device=dev)
    rows = []

    def row(name, ms_, units, unit, algo_bytes, extra=None):
        r = {"row": name, "ms_per_step": ms_, "value": units / (ms_ * 1e-3), "unit": unit,
             "roofline": {"bound": "hbm", "achieved": algo_bytes / (ms_ * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                          "frac": algo_bytes / (ms_ * 1e-3) / 1e9 / peak, "peak_source": kind, "algorithmic_bytes": algo_bytes}}
        if extra:
            r.update(extra)
        rows.append(r)
        print(json.dumps(r), flush=True)

    # ---- f-1 NeMo frontend
    cfg = ms.BatchLogMelConfig(n_mels=128, preemphasis=0.97, normalize_per_feature=True)   # Parakeet-style
    nemo = ms.BatchLogMelSpectrogram(cfg)
    cols = nemo.padded_frames(n)
    F = nemo.num_frames(n)
    out = torch.empty((clips, 128, cols), dtype=torch.float32, device=dev)
    t = timeit(lambda: nemo.compute_device(x, clips, n, n, out, layout=ms.LAYOUT_MEL_MAJOR, stream=st), args.steps, st)
    row("f-1 NeMo BatchLogMel 128 mel n_fft 512 (+ per-feature normalisation), 1024 x 10 s", t, clips * F, "frames/s",
        clips * (4 * n + 4 * 128 * cols), {"frames_per_clip": F, "padded_cols": cols,
                                            "note": "normalisation kernel re-reads and rewrites the features (counted once here)"})
    del out
    # ---- f-3a interleaved store
    mel = ms.CudaMelSpectrogram(400, 160, 16000.0, 80)
    Fw = mel.num_frames(n)
    W = 1000
    img = torch.empty((clips, 80, W), dtype=torch.float32, device=dev)
    t = timeit(lambda: mel.compute_interleaved_device(x, clips, n, n, W, img, stream=st), args.steps, st)
    row("f-3a fused kernel -> interleave_frames image (80 x 1000, min_width 1000), 1024 x 10 s", t, clips * Fw, "frames/s",
        clips * (4 * n + 4 * 80 * W))
    # ---- f-3b quantise / dequantise
    tga = torch.empty((clips, 26 + 80 * W), dtype=torch.uint8, device=dev)
    px = clips * 80 * W
    t = timeit(lambda: mel.quantize_tga_device(img, clips, 80, W, tga, stream=st), args.steps, st)
    row("f-3b TGA quantise (min/max + quantise), 1024 images 80 x 1000", t, px, "pixels/s", px * 9)
    back = torch.empty_like(img)
    t = timeit(lambda: mel.dequantize_tga_device(tga, clips, 80, W, back, stream=st), args.steps, st)
    row("f-3b TGA dequantise, 1024 images 80 x 1000", t, px, "pixels/s", px * 5)
    # ---- f-4 VAD
    sm = torch.empty((clips, W - 2), dtype=torch.uint8, device=dev)
    rw = torch.empty((clips, W - 2), dtype=torch.uint8, device=dev)
    act = torch.empty((clips, W, 3), dtype=torch.int32, device=dev)
    ds = ms.DetectionSettings()

    def vad():
        mel.vad_boundaries_device(img, clips, 80, W, ds, sm, d_raw=rw, stream=st)
        mel.vad_activity_device(rw, clips, 80, W, ds, act, stream=st)
    t = timeit(vad, args.steps, st)
    row("f-4 VAD boundaries + per-frame activity (f64 Sobel), 1024 images 80 x 1000", t, clips * W, "frames/s",
        px * 4 + clips * (W - 2) * 2 + clips * W * 12)
    # ---- general plan (melspec_generic.cuh): sizes outside the two specialised kernels
    del img, tga, back, sm, rw, act
    for fft, hop, nm in ((1024, 256, 128), (2048, 512, 80), (480, 160, 80)):
        g = ms.CudaMelSpectrogram(fft, hop, 16000.0, nm)
        Fg = g.num_frames(n)
        og = torch.empty((clips, Fg, nm), dtype=torch.float32, device=dev)
        t = timeit(lambda: g.compute_device(x, clips, n, n, og, stream=st), args.steps, st)
        row(f"general plan: Whisper fft {fft} hop {hop} {nm} mel, 1024 x 10 s", t, clips * Fg, "frames/s",
            clips * (4 * n + 4 * nm * Fg), {"frames_per_clip": Fg})
        g.close()
        del og
    # ---- plan 512 in its three modes at the cfg2 / cfg3 size (where the Kaldi prologue and CMN cost show)
    for name, mk, nm in (("plan 400: Whisper large-v3 style, fft 400 hop 160 128 mel", lambda: ms.CudaMelSpectrogram(400, 160, 16000.0, 128), 128),
                         ("plan 512: Whisper fft 512 hop 160 80 mel (golden-file configuration)", lambda: ms.CudaMelSpectrogram(512, 160, 16000.0, 80), 80),
                         ("plan 512: Kaldi fbank 80 bins, CMN off", lambda: ms.Fbank(ms.FbankConfig(apply_cmn=False)), 80),
                         ("plan 512: Kaldi fbank 80 bins + CMN (cfg3)", lambda: ms.Fbank(ms.FbankConfig()), 80)):
        hh = mk()
        Fh = hh.num_frames(n)
        oh = torch.empty((clips, Fh, nm), dtype=torch.float32, device=dev)
        t = timeit(lambda: hh.compute_device(x, clips, n, n, oh, stream=st), args.steps, st)
        row(name + ", 1024 x 10 s", t, clips * Fh, "frames/s", clips * (4 * n + 4 * nm * Fh), {"frames_per_clip": Fh})
        hh.close()
        del oh
    fb8 = ms.Fbank(ms.FbankConfig(sample_rate=8000.0, num_mel_bins=40))
    Fk = fb8.num_frames(n)
    ok = torch.empty((clips, Fk, 40), dtype=torch.float32, device=dev)
    t = timeit(lambda: fb8.compute_device(x, clips, n, n, ok, stream=st), args.steps, st)
    row("general plan: Kaldi fbank 8 kHz (200-sample frames, fft 256, shift 80) 40 bins + CMN, 1024 x 20 s", t, clips * Fk, "frames/s",
        clips * (4 * n + 4 * 40 * Fk), {"frames_per_clip": Fk})
    return rows


if __name__ == "__main__":
    main()
