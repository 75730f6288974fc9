// Links the prebuilt CUDA library (built by `make -C zkstark_b200/csrc`, nvcc -arch=sm_100a).
// Headers are hand-written in src/ffi.rs from include/zkstark_b200.h (no bindgen binary in the image).
fn main() {
    let dir = std::env::var("ZKSTARK_B200_LIB_DIR").unwrap_or_else(|_| "../../zkstark_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=zkstark_b200");
    println!("cargo:rerun-if-env-changed=ZKSTARK_B200_LIB_DIR");
}
