// Locates libpfgpu.so: PFGPU_LIB_DIR, else the in-tree build (rust_robotics_b200/libpfgpu.so, produced by
// `python -c "import __graft_entry__ as g; g.build()"` or `python rust_robotics_b200/build.py`).
use std::{env, path::PathBuf};

fn main() {
    let dir = env::var("PFGPU_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..")       // rust/pfgpu-sys -> rust_robotics_b200/
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=pfgpu");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=PFGPU_LIB_DIR");
}
