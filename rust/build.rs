// build.rs -- compiles the CUDA sources with nvcc for sm_100a and links them (UNCOMPILED, see README.md).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from("frieda_b200/csrc");
    let lib = out.join("libfrieda_b200.so");
    let mut cmd = Command::new("nvcc");
    cmd.args([
        "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
        "--expt-relaxed-constexpr", "-shared", "-x", "cu", "-o",
    ])
    .arg(&lib);
    for f in ["ctx.cu", "lde.cu", "merkle.cu", "fri.cu", "decommit.cu", "verify_batch.cu", "proof.cpp", "verify.cpp"] {
        cmd.arg(csrc.join(f));
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    assert!(cmd.status().expect("nvcc not found").success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=frieda_b200");
}
