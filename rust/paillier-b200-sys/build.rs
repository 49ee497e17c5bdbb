// Links against the prebuilt CUDA library (python -m paillier_halo2_b200.build).  PB200_LIB_DIR points at the directory that
// holds libpaillier_b200.so (default: ../../paillier_halo2_b200 relative to this crate).
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("PB200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../paillier_halo2_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=paillier_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=PB200_LIB_DIR");
}
