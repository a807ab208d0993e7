// build.rs of the reference crate: link the B200 library.  ZKB200_LIB_DIR = directory holding libzkb200.so
// (zksnark-rs_b200/ in the zkb200 repository after `python zksnark-rs_b200/build.py`).
fn main() {
    let dir = std::env::var("ZKB200_LIB_DIR").expect("set ZKB200_LIB_DIR to the directory of libzkb200.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=zkb200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=ZKB200_LIB_DIR");
}
