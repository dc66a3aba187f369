#!/usr/bin/env python
"""Regenerate tests/golden/ from the reference's own test fixtures.

Run in the build container (where /root/reference is mounted read-only):

    python tests/golden/make_fixtures.py

/root/reference does not exist on the GPU box, so the *data* the reference's tests pin
(`src/rb.rs:134-179`, `src/mel.rs:837-871`, `src/fbank.rs:439-535`) is copied here as
small arrays.  Nothing in here is reference source code; these are the reference's golden
vectors and the audio clip they were computed from.

  jfk_pcm_f32.npy        176000 f32 samples = payload of testdata/jfk_f32le.wav `data` chunk
                         (RIFF chunk walk, the way src/fbank.rs:324-352 does it)
  rust_jfk_golden.npy    (80,1097) f32 — stream path, fft 512 / hop 160 / 80 mel (src/rb.rs:136-141)
  mel_filters_80x201.npy (80,201) f32 — testdata/mel_filters.npz['mel_80']  (src/mel.rs:837-850)
  nemo_filters_80x257.npy(80,257) f32 — testdata/nemo_mel_filters.npz['banks'][0] (src/mel.rs:852-871)
  kaldi_fbank_jfk.npy    (80,1098) f32 — testdata/kaldi_native_fbank_jfk.npz['features'] (src/fbank.rs:439-535)
  quantized_mel_golden.tga  80 x 1100 8-bit TGA (src/quant.rs format) — testdata/quantized_mel_golden.tga, the fixture of the
                         reference's VAD tests (src/vad.rs:712,742; tests/vad_regression.rs:157,215).  Columns 2..1099 are the
                         quantised Whisper fft-400 / hop-160 stream mel of the JFK clip (verified bit-exact against the oracle).
  vad/blank/*.tga, vad/speech/*.tga   the 7 + 5 quantised mel images of the reference's VAD known-answer test
                         (src/vad.rs:621-668: vad_on must be false on blank/, true on speech/)
"""
import os
import struct
import sys

import numpy as np

REF = os.environ.get("MELSPEC_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def read_wav_f32(path):
    raw = open(path, "rb").read()
    assert raw[:4] == b"RIFF" and raw[8:12] == b"WAVE"
    pos = 12
    fmt = None
    while pos + 8 <= len(raw):
        cid = raw[pos:pos + 4]
        size = struct.unpack("<I", raw[pos + 4:pos + 8])[0]
        body = pos + 8
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", raw[body:body + 16])
        elif cid == b"data":
            assert fmt is not None and fmt[1] == 1 and fmt[2] == 16000 and fmt[5] == 32
            return np.frombuffer(raw[body:body + size], dtype="<f4").copy(), body
        pos = body + size + (size & 1)
    raise ValueError("no data chunk")


def main():
    td = os.path.join(REF, "testdata")
    if not os.path.isdir(td):
        sys.exit(f"{td} not found (run this in the build container)")
    pcm, off = read_wav_f32(os.path.join(td, "jfk_f32le.wav"))
    assert pcm.shape == (176000,) and off == 114, (pcm.shape, off)
    np.save(os.path.join(HERE, "jfk_pcm_f32.npy"), pcm)
    np.save(os.path.join(HERE, "rust_jfk_golden.npy"), np.load(os.path.join(td, "rust_jfk_golden.npy")))
    np.save(os.path.join(HERE, "mel_filters_80x201.npy"), np.load(os.path.join(td, "mel_filters.npz"))["mel_80"])
    np.save(os.path.join(HERE, "nemo_filters_80x257.npy"), np.load(os.path.join(td, "nemo_mel_filters.npz"))["banks"][0])
    np.save(os.path.join(HERE, "kaldi_fbank_jfk.npy"), np.load(os.path.join(td, "kaldi_native_fbank_jfk.npz"))["features"])
    with open(os.path.join(td, "quantized_mel_golden.tga"), "rb") as src, open(os.path.join(HERE, "quantized_mel_golden.tga"), "wb") as dst:
        dst.write(src.read())
    for kind in ("blank", "speech"):            # fixtures of src/vad.rs:621-668 (test_speech_detection)
        os.makedirs(os.path.join(HERE, "vad", kind), exist_ok=True)
        for f in sorted(os.listdir(os.path.join(td, kind))):
            if f.endswith(".tga"):
                with open(os.path.join(td, kind, f), "rb") as src, open(os.path.join(HERE, "vad", kind, f), "wb") as dst:
                    dst.write(src.read())
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npy"):
            a = np.load(os.path.join(HERE, f))
            print(f, a.shape, a.dtype)


if __name__ == "__main__":
    main()
