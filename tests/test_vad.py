"""VAD over the mel image: Sobel edge count + majority smoothing, reference src/vad.rs:251-486, and the
per-frame activity of VoiceActivityDetector::add_activity (src/vad.rs:163-207).  CPU tests restate the reference's own
known-answer test (src/vad.rs:621-668) on its fixtures; GPU tests require the device masks to equal the oracle's exactly."""
import glob
import os

import numpy as np
import pytest

import melspec_oracle as o


def _images(golden_dir, kind):
    for f in sorted(glob.glob(os.path.join(golden_dir, "vad", kind, "*.tga"))):
        yield os.path.basename(f), o.parse_tga_8bit(open(f, "rb").read()).reshape(80, -1)


# ------------------------------------------------------------------------------------------------ CPU: oracle pinned
def test_oracle_speech_detection_known_answers(golden_dir):
    # src/vad.rs:621-668: settings (min_energy 1.0, min_y 10, min_x 10, min_mel 0); blank -> vad_on false, speech -> true
    n = 0
    for name, img in _images(golden_dir, "blank"):
        _, inter = o.vad_boundaries(img, 1.0, 10, 10, 0)
        assert o.vad_on(inter, 10) is False, name
        n += 1
    for name, img in _images(golden_dir, "speech"):
        _, inter = o.vad_boundaries(img, 1.0, 10, 10, 0)
        assert o.vad_on(inter, 10) is True, name
        n += 1
    assert n == 12


def test_oracle_vad_small_cases():
    assert o.vad_boundaries(np.zeros((2, 10))) == ([], [])                # src/vad.rs:264-266
    assert o.vad_boundaries(np.zeros((80, 2))) == ([], [])
    non, inter = o.vad_boundaries(np.zeros((80, 10)), min_y=0)             # min_y == 0: every column is active (382-385)
    assert inter == list(range(8)) and non == []
    assert o.vad_smooth_mask([1, 0, 0, 0, 0, 0, 0, 0, 0, 0], 4).tolist() == [False] * 10
    assert o.vad_smooth_mask([1, 1, 1, 0, 0, 0], 4).tolist() == [True] * 5 + [False]   # 3 of 5, 3 of 6 (at least half), 2 of 5
    assert o.vad_on([5], 1) is False and o.vad_on([5, 6], 1) is True and o.vad_on([], 1) is False   # src/vad.rs:226-249
    assert o.vad_leading_active_columns([0, 1, 2, 5]) == 3 and o.vad_leading_active_columns([1, 2]) == 0
    # an edge: a bright block produces gradients on its border
    img = np.zeros((80, 40)); img[20:60, 15:25] = 3.0
    raw = o.vad_raw_classification(img, 0.98, 11, 2)
    assert np.flatnonzero(raw).tolist() == [13, 14, 23, 24]           # patches x..x+2 that straddle the block's two vertical borders


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def mel400():
    import mel_spec_b200 as ms
    ms.build()
    h = ms.CudaMelSpectrogram(400, 160, 16000.0, 80)
    yield h
    h.close()


@pytest.mark.gpu
def test_gpu_vad_matches_oracle_on_reference_fixtures(mel400, golden_dir):
    import mel_spec_b200 as ms
    st = ms.DetectionSettings(1.0, 10, 10, 0)
    for kind, expect in (("blank", False), ("speech", True)):
        for name, img in _images(golden_dir, kind):
            ei = mel400.vad_boundaries(img, st)
            non, inter = o.vad_boundaries(img, 1.0, 10, 10, 0)
            assert ei.intersected() == inter and ei.non_intersected() == non, name
            assert ms.vad_on(ei, 10) is expect, name
    # default settings and a min_y == 0 run on the golden JFK image, plus the TGA round trip through the device parser
    raw = open(os.path.join(golden_dir, "quantized_mel_golden.tga"), "rb").read()
    img = mel400.parse_tga_8bit(raw).reshape(80, 1100)
    for st in (ms.DetectionSettings(), ms.DetectionSettings(1.0, 3, 6, 0), ms.DetectionSettings(0.5, 0, 5, 2)):
        ei = mel400.vad_boundaries(img, st)
        non, inter = o.vad_boundaries(img, st.min_energy, st.min_y, st.min_x, st.min_mel)
        assert ei.intersected() == inter and ei.non_intersected() == non
        assert len(inter) + len(non) == 1098


@pytest.mark.gpu
def test_gpu_vad_activity_stream(mel400, golden_dir):
    import mel_spec_b200 as ms
    raw = open(os.path.join(golden_dir, "quantized_mel_golden.tga"), "rb").read()
    img = o.parse_tga_8bit(raw).reshape(80, 1100)[:, :400]
    for st in (ms.DetectionSettings(), ms.DetectionSettings(1.0, 3, 12, 0), ms.DetectionSettings(1.0, 3, 2, 0)):
        got = mel400.vad_activities(img, st, ms.VadFrameTiming(400, 160, 16000.0))
        want = o.vad_activity_stream(img, st.min_energy, st.min_y, st.min_x, st.min_mel)
        assert len(got) == len(want) == 400 - st.min_x + 1
        for g, w in zip(got, want):
            assert (g.frame_index, g.active, g.leading_active_columns, g.active_columns, g.window_columns) == w
        assert got[0].timestamps.start_ms == (st.min_x - 1) * 10 and got[0].timestamps.end_ms == (st.min_x - 1) * 10 + 25


@pytest.mark.gpu
def test_gpu_vad_on_kernel_output_batch(mel400, jfk):
    """PCM -> fused kernel (interleaved image) -> VAD kernel, batch of clips, all on the device; the masks must equal the
    oracle's VAD applied to the kernel's own f32 image, and speech must be found in the JFK clip."""
    import torch
    import mel_spec_b200 as ms
    pcm = np.stack([jfk[:160000], jfk[16000:176000], np.zeros(160000, np.float32)])
    x = torch.from_numpy(pcm).cuda()
    w = 998
    img = torch.empty((3, 80, w), dtype=torch.float32, device="cuda")
    mel400.compute_interleaved_device(x, 3, 160000, 160000, 0, img)
    st = ms.DetectionSettings()
    rawm = torch.zeros((3, w - 2), dtype=torch.uint8, device="cuda")
    sm = torch.zeros((3, w - 2), dtype=torch.uint8, device="cuda")
    act = torch.zeros((3, w, 3), dtype=torch.int32, device="cuda")
    mel400.vad_boundaries_device(img, 3, 80, w, st, sm, d_raw=rawm)
    mel400.vad_activity_device(rawm, 3, 80, w, st, act)
    torch.cuda.synchronize()
    im, s, a = img.cpu().numpy(), sm.cpu().numpy(), act.cpu().numpy()
    for i in range(3):
        non, inter = o.vad_boundaries(im[i], st.min_energy, st.min_y, st.min_x, st.min_mel)
        assert np.flatnonzero(s[i]).tolist() == inter
        want = o.vad_activity_stream(im[i], st.min_energy, st.min_y, st.min_x, st.min_mel)
        assert [(k, bool(a[i, k, 0]), int(a[i, k, 1]), int(a[i, k, 2])) for k in range(st.min_x - 1, w)] == [t[:4] for t in want]
        assert (a[i, :st.min_x - 1] == -1).all()
    assert s[0].sum() > 100 and s[2].sum() == 0


@pytest.mark.gpu
def test_gpu_vad_decisions_at_the_threshold(mel400):
    """The VAD kernel evaluates the reference's f64 expression in the reference's operation order (src/vad.rs:284-316).  Images whose
    Sobel energies sit within a few ulp of min_energy^2 — exact ties included — must give the oracle's masks, for single-pixel
    decisions (min_y = 1).  (Round 2 tried an fp32 evaluation with a rigorous error band and f64 only inside the band: identical masks
    on these cases, but 0.297 instead of 0.209 ms per 1024 images — the kernel is not bound by the FP64 pipe on B200 — so it was dropped.)"""
    import mel_spec_b200 as ms
    rng = np.random.default_rng(17)
    h, w = 24, 300
    for s0 in (0.125, 0.3, 1.0 / 3.0, 7.77):
        # a horizontal ramp whose step varies by a few ulp from column to column: gx = 4 (s_x + s_x+1) ~ 8 s0, gy ~ 0
        steps = np.float32(s0) * (1.0 + rng.integers(-3, 4, w).astype(np.float64) * 2.0 ** -23)
        row = np.cumsum(steps).astype(np.float32)
        img = np.tile(row, (h, 1))
        img[rng.integers(0, h, 40), rng.integers(0, w, 40)] += np.float32(s0 * 2.0 ** -22)      # a little vertical structure
        for min_energy in (8.0 * s0, float(np.float32(8.0 * s0)), 8.0 * s0 * (1 + 2.0 ** -24), 8.0 * s0 * (1 - 2.0 ** -24)):
            for min_y in (1, 3):
                st = ms.DetectionSettings(min_energy, min_y, 4, 0)
                ei = mel400.vad_boundaries(img, st)
                non, inter = o.vad_boundaries(img, min_energy, min_y, 4, 0)
                assert ei.intersected() == inter and ei.non_intersected() == non, (s0, min_energy, min_y)
    # random images, thresholds equal to the f64 energy of actual pixels (ties) and to their neighbours in f64
    img = rng.standard_normal((40, 200)).astype(np.float32)
    d = img.astype(np.float64)
    gx = (d[:-2, 2:] + 2 * d[1:-1, 2:] + d[2:, 2:]) - (d[:-2, :-2] + 2 * d[1:-1, :-2] + d[2:, :-2])
    gy = (d[2:, :-2] + 2 * d[2:, 1:-1] + d[2:, 2:]) - (d[:-2, :-2] + 2 * d[:-2, 1:-1] + d[:-2, 2:])
    e = gx * gx + gy * gy
    for k in rng.integers(0, e.size, 6):
        thr = float(np.sqrt(e.reshape(-1)[k]))
        for me in (thr, np.nextafter(thr, 0.0), np.nextafter(thr, 1e9)):
            st = ms.DetectionSettings(float(me), 1, 3, 0)
            ei = mel400.vad_boundaries(img, st)
            non, inter = o.vad_boundaries(img, float(me), 1, 3, 0)
            assert ei.intersected() == inter and ei.non_intersected() == non
    # non-finite pixels fall back to the f64 path (NaN energies compare false, like the reference's)
    img[5, 7] = np.inf
    img[9, 100] = np.nan
    st = ms.DetectionSettings(1.0, 2, 3, 0)
    ei = mel400.vad_boundaries(img, st)
    non, inter = o.vad_boundaries(img, 1.0, 2, 3, 0)
    assert ei.intersected() == inter and ei.non_intersected() == non
