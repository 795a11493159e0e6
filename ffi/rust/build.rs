// Link lines for libjne.so (built by nvcc: `python -m johansen_null_eigenspectra_b200.build`).
// Nothing is compiled by cargo; there is no CPU fallback to fall back to when the library is missing.
fn main() {
    let dir = std::env::var("JNE_LIB_DIR").expect("set JNE_LIB_DIR to the directory holding libjne.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=jne");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=JNE_LIB_DIR");
}
