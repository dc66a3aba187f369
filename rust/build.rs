// Replacement for the reference's build.rs (which compiles src/cuda_kernels.cu with nvcc and links cufft + cudart,
// reference build.rs:5-50): the CUDA code now lives in the prebuilt libmelspec_b200.so, so the build script only has
// to tell cargo where that library is.  SOURCE ONLY (no cargo in this image).
use std::path::PathBuf;

fn main() {
    if std::env::var("CARGO_FEATURE_CUDA").is_err() {
        return;
    }
    // default: <repo>/mel-spec_b200/lib next to this crate, as an absolute path (cargo runs build scripts from its own cwd)
    let dir = std::env::var("MELSPEC_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(std::env::var("CARGO_MANIFEST_DIR").expect("cargo sets CARGO_MANIFEST_DIR")).join("..").join("mel-spec_b200").join("lib")
    });
    let dir = dir.canonicalize().unwrap_or(dir);
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=melspec_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=MELSPEC_B200_LIB_DIR");
}
