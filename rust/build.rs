// Link against libblsgpu.so.  BLSGPU_LIB_DIR = directory holding the library built by `python __graft_entry__.py`
// (bls_verify_gadget_b200/ in the repo); the CUDA runtime it needs is found through its own RUNPATH / LD_LIBRARY_PATH.
fn main() {
    let dir = std::env::var("BLSGPU_LIB_DIR").unwrap_or_else(|_| "../bls_verify_gadget_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=blsgpu");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=BLSGPU_LIB_DIR");
}
