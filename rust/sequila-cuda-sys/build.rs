// UNCOMPILED — see rust/README.md.
// Links the in-tree shared library; SEQUILA_CUDA_LIB_DIR points at the directory holding libsequila_cuda.so
// (sequila_native_b200/ of this repository after `make -C sequila_native_b200/csrc`).
fn main() {
    let dir = std::env::var("SEQUILA_CUDA_LIB_DIR").expect("set SEQUILA_CUDA_LIB_DIR to the directory of libsequila_cuda.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=sequila_cuda");
    println!("cargo:rerun-if-env-changed=SEQUILA_CUDA_LIB_DIR");
}
