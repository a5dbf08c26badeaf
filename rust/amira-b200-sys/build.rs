//! Builds libamira_b200.so with nvcc for sm_100a and links it.
//! Replaces the CUDA half of the reference's build script (build.rs:11-114: `nvcc -c src/cuda/cuda_helper.cu`, the
//! `cargo:rustc-link-arg=cuda_helper.o` line and every Triton link line — none of them is needed on this path).
//! The same source list and flags as amira-rust-asr-server_b200/build.py, which is what the Python tests build.
use std::path::PathBuf;
use std::process::Command;

/// Translation units of the library (amira-rust-asr-server_b200/csrc); tests/test_rust_binding.py keeps this list equal to
/// build.py's SOURCES.
const SOURCES: &[&str] = &[
    "api.cu",
    "frontend.cu",
    "decoder.cu",
    "decoder_tc.cu",
    "decoder_ws.cu",
    "tables.cpp",
    "host_pipeline.cpp",
    "host_stream.cpp",
    "host_wire.cpp",
];

fn main() {
    // AMIRA_B200_ROOT = a checkout of the library repository (contains include/ and amira-rust-asr-server_b200/csrc/)
    let root = PathBuf::from(std::env::var("AMIRA_B200_ROOT").unwrap_or_else(|_| "../..".to_string()));
    let csrc = root.join("amira-rust-asr-server_b200").join("csrc");
    let include = root.join("include");
    let out = PathBuf::from(std::env::var("OUT_DIR").expect("OUT_DIR"));
    let lib = out.join("libamira_b200.so");

    println!("cargo:rerun-if-changed=build.rs");
    println!("cargo:rerun-if-env-changed=AMIRA_B200_ROOT");
    println!("cargo:rerun-if-changed={}", include.join("amira_b200.h").display());
    for s in SOURCES {
        println!("cargo:rerun-if-changed={}", csrc.join(s).display());
    }

    // A prebuilt library can be used instead (AMIRA_B200_LIB_DIR), e.g. the one build.py left in the checkout.
    if let Ok(dir) = std::env::var("AMIRA_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=amira_b200");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
        return;
    }

    let nvcc = std::env::var("NVCC").unwrap_or_else(|_| "nvcc".to_string());
    let mut cmd = Command::new(&nvcc);
    cmd.args([
        "-gencode",
        "arch=compute_100a,code=sm_100a", // B200 only: no multi-arch fatbin, no fallback path
        "-lineinfo",
        "-O3",
        "-std=c++17",
        "-shared",
        "-Xcompiler",
        "-fPIC",
    ]);
    cmd.arg("-I").arg(&include).arg("-I").arg(&csrc);
    for s in SOURCES {
        cmd.arg(csrc.join(s));
    }
    cmd.arg("-o").arg(&lib).arg("-lcudart");
    let status = cmd.status().unwrap_or_else(|e| panic!("cannot run {nvcc}: {e}"));
    assert!(status.success(), "nvcc failed building libamira_b200.so");

    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=amira_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", out.display());
    // CUDA runtime search paths, as in the reference's build.rs:100-107
    println!("cargo:rustc-link-search=native=/usr/local/cuda/targets/x86_64-linux/lib");
    println!("cargo:rustc-link-lib=cudart");
    println!("cargo:rustc-link-lib=stdc++");
}
