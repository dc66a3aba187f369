// Replacement for the reference's build.rs (which compiles src/cuda_kernels.cu with nvcc and links cufft + cudart,
// reference build.rs:5-50): the CUDA code now lives in the prebuilt libmelspec_b200.so, so the build script only has
// to tell cargo where that library is.  SOURCE ONLY (no cargo in this image).
fn main() {
    if std::env::var("CARGO_FEATURE_CUDA").is_err() {
        return;
    }
    let dir = std::env::var("MELSPEC_B200_LIB_DIR").unwrap_or_else(|_| "../mel-spec_b200/lib".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=melspec_b200");
    println!("cargo:rerun-if-env-changed=MELSPEC_B200_LIB_DIR");
}
