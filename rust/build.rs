// build.rs -- compiles the CUDA sources with nvcc for sm_100a, the host-only SIMD sources with the host
// compiler, and links them (UNCOMPILED, see README.md).  Mirrors frieda_b200/build.py.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from("frieda_b200/csrc");
    let lib = out.join("libfrieda_b200.so");
    // host-only objects: the verifier's 8-/16-way hashing (entered after a runtime CPU check) and that check
    let mut host_objs = Vec::new();
    for (f, isa) in [("blake2s_x8.cpp", Some("-mavx2")), ("blake2s_x16.cpp", Some("-mavx512f")), ("cpu_features.cpp", None)] {
        let obj = out.join(format!("{f}.o"));
        let mut cxx = Command::new(env::var("CXX").unwrap_or_else(|_| "g++".into()));
        cxx.args(["-O3", "-std=c++17", "-fPIC", "-c"]).arg(csrc.join(f)).arg("-o").arg(&obj);
        if let Some(flag) = isa {
            cxx.arg(flag);
        }
        assert!(cxx.status().expect("host compiler not found").success(), "host compile failed");
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
        host_objs.push(obj);
    }
    let mut cmd = Command::new("nvcc");
    cmd.args([
        "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
        "--expt-relaxed-constexpr", "-shared", "-x", "cu", "-o",
    ])
    .arg(&lib);
    for f in ["ctx.cu", "lde.cu", "merkle.cu", "fri.cu", "decommit.cu", "verify_batch.cu", "proof.cpp", "verify.cpp", "split_proof.cpp"] {
        cmd.arg(csrc.join(f));
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    cmd.args(["-x", "none"]); // the objects that follow are not CUDA sources
    for o in &host_objs {
        cmd.arg(o);
    }
    assert!(cmd.status().expect("nvcc not found").success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=frieda_b200");
}
