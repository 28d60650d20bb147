// Compiles the hand-written kernels with nvcc for sm_100a and links libparticular_cuda.so.
// There is no wgpu, no multi-backend dispatch and no CPU fallback: without nvcc the build fails.
// `PARTICULAR_CUDA_CSRC` may point at the kernel sources (default: ../../particular_b200/csrc, the
// layout of the repository this crate ships in).  NCCL is dlopen()ed at run time by comm.cu, so
// there is no link-time NCCL dependency.
use std::{env, path::PathBuf, process::Command};

const UNITS: [&str; 11] = [
    "context.cu", "bruteforce.cu", "barneshut.cu", "bh_build.cu", "bh_radix_build.cu", "bh_traverse.cu",
    "bh_multigpu.cu", "comm.cu", "sim.cu", "custom.cu", "probe.cu",
];

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = env::var("PARTICULAR_CUDA_CSRC")
        .map(PathBuf::from)
        .unwrap_or_else(|_| PathBuf::from("../../particular_b200/csrc"));
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let mut objs = Vec::new();
    for unit in UNITS {
        let obj = out.join(unit).with_extension("o");
        let status = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17"])
            .args(["-Xcompiler", "-fPIC", "-c"])
            .arg(csrc.join(unit))
            .arg("-o")
            .arg(&obj)
            .status()
            .expect("nvcc not found: particular-cuda has no CPU fallback");
        assert!(status.success(), "nvcc failed for {unit}");
        println!("cargo:rerun-if-changed={}", csrc.join(unit).display());
        objs.push(obj);
    }
    for header in ["common.cuh", "ptx.cuh", "bh.cuh"] {
        println!("cargo:rerun-if-changed={}", csrc.join(header).display());
    }
    let lib = out.join("libparticular_cuda.so");
    let status = Command::new(&nvcc)
        .arg("-shared")
        .arg("-o")
        .arg(&lib)
        .args(&objs)
        .arg("-ldl")
        .status()
        .expect("nvcc link step");
    assert!(status.success(), "linking libparticular_cuda.so failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=particular_cuda");
}
