// UNVERIFIED (no rustc in the build image).
// Links the prebuilt libb200sdf.so (make -C versatiles_glyphs_rs_b200/csrc; nvcc -gencode arch=compute_100a,code=sm_100a).
// B200SDF_LIB_DIR = directory holding libb200sdf.so (default: ../../versatiles_glyphs_rs_b200 relative to this crate).
use std::{env, path::PathBuf};

fn main() {
    let dir = env::var("B200SDF_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../versatiles_glyphs_rs_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=b200sdf");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=B200SDF_LIB_DIR");
    println!("cargo:rerun-if-changed=../../include/b200sdf.h");
}
