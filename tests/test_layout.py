"""Layout of the C ABI's POD structs, checked three ways (CPU only): a compiled C probe prints offsetof / sizeof of every
field of `melspec_config` and `melspec_vad_settings` (tests/cpp/layout_probe.c); the ctypes mirror (mel-spec_b200/_lib.py) and
the `#[repr(C)]` structs of the Rust shim (rust/src/ffi.rs, `// offset N` comments + `const _` assertions) must agree with it
field by field, in order.  A reordered or retyped field would otherwise pass every functional test."""
import ctypes as C
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _probe():
    exe = os.path.join(ROOT, "build", "layout_probe")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "layout_probe.c"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    rows = [ln.split() for ln in out.strip().splitlines()]
    return [(r[0], int(r[1]), int(r[2])) for r in rows]


def _rust_struct(name):
    src = open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()
    body = re.search(r"pub struct %s \{(.*?)\n\} // sizeof (\d+)" % name, src, re.S)
    assert body, f"struct {name} not found in rust/src/ffi.rs"
    fields = re.findall(r"pub (\w+): (\w+),\s*// offset (\d+)", body.group(1))
    return [(f, t, int(o)) for f, t, o in fields], int(body.group(2))


def test_struct_layouts_agree_between_c_ctypes_and_rust():
    from mel_spec_b200._lib import MelspecConfig, VadSettings
    probe = _probe()
    rust_size = {"i32": 4, "f64": 8}
    for cname, ctype, rname in (("melspec_config", MelspecConfig, "MelspecConfig"), ("melspec_vad_settings", VadSettings, "VadSettings")):
        c_fields = [(n.split(".")[1], off, sz) for n, off, sz in probe if n.startswith(cname + ".") and not n.endswith(".sizeof")]
        c_sizeof = next(off for n, off, _ in probe if n == cname + ".sizeof")
        # ctypes: same names, same order, same offsets and sizes
        assert [f for f, _ in ctype._fields_] == [f for f, _, _ in c_fields]
        for f, off, sz in c_fields:
            d = getattr(ctype, f)
            assert (d.offset, d.size) == (off, sz), (cname, f, d.offset, d.size, off, sz)
        assert C.sizeof(ctype) == c_sizeof
        # Rust: same names, same order, offsets from the comments, sizes from the types
        r_fields, r_sizeof = _rust_struct(rname)
        assert [f for f, _, _ in r_fields] == [f for f, _, _ in c_fields]
        for (f, t, off), (_, coff, csz) in zip(r_fields, c_fields):
            assert (off, rust_size[t]) == (coff, csz), (rname, f, off, t, coff, csz)
        assert r_sizeof == c_sizeof
    abi = next(off for n, off, _ in probe if n == "abi")
    assert f"pub const ABI_VERSION: i32 = {abi};" in open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()


def test_rust_extern_block_declares_only_exported_symbols():
    """Every `pub fn melspec_*` of the Rust extern block exists in the header with the same number of parameters."""
    src = open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "melspec_b200.h")).read(), flags=re.S)
    block = src[src.index('unsafe extern "C" {'):]
    for name, args in re.findall(r"pub fn (melspec_\w+)\((.*?)\)", block, re.S):
        m = re.search(r"\b%s\s*\((.*?)\)\s*;" % name, hdr, re.S)
        assert m, f"{name} is bound in rust/src/ffi.rs but not declared in include/melspec_b200.h"
        n_rust = 0 if not args.strip() else len([a for a in args.split(",") if a.strip()])
        c_args = m.group(1).strip()
        n_c = 0 if c_args in ("", "void") else len([a for a in c_args.split(",") if a.strip()])
        assert n_rust == n_c, (name, n_rust, n_c)
