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
