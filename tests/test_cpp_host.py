"""The compiled-language host side above the C ABI: include/melspec_b200.hpp (C++17 mirror of the reference's Rust prelude)
driven by tests/cpp/host_mirror_test.cpp, which restates the reference's own test expectations (src/rb.rs:134-179,
src/cuda.rs:488-545, tests/readme_examples.rs, src/fbank.rs:439-535, src/mel.rs:943-961, src/vad.rs:621-668)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "host_mirror_test")


def _build():
    import mel_spec_b200 as ms
    ms.build()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    libdir = os.path.dirname(ms.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"), "-L", libdir, "-lmelspec_b200",
                           f"-Wl,-rpath,{libdir}", "-o", EXE])


def test_cpp_host_mirror_builds_and_device_free_checks(golden_dir):
    _build()
    out = subprocess.run([EXE, golden_dir, "nogpu"], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout + out.stderr


@pytest.mark.gpu
def test_cpp_host_mirror_reference_expectations(golden_dir):
    _build()
    out = subprocess.run([EXE, golden_dir], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout + out.stderr
    assert "FAIL" not in out.stdout
