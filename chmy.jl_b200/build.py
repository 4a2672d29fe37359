"""Builds libchmy_b200.so (hand-written CUDA for sm_100a + the C ABI of include/chmy_b200.h) in-tree with nvcc."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libchmy_b200.so")
SOURCES = ["api.cu", "ops.cu", "ops_fast.cu", "ops_fast2d.cu", "ops_fused.cu", "ops_fused2d.cu", "bc.cu", "comm.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",              # exact arithmetic contract: never contract a*b+c
    "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC,-O2,-Wall",
    "-Xptxas", "-v",
    "-shared",
]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "chmy_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    env = dict(os.environ)
    env.pop("CC", None)       # the image exports CC/CXX wrappers; let nvcc pick the distro g++
    env.pop("CXX", None)
    cmd = [nvcc] + NVCC_FLAGS + ["-ccbin", "/usr/bin/g++", "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libchmy_b200.so")
    logdir = os.path.join(HERE, "build")           # git-ignored: ptxas -v output (registers, spills) of the last build
    os.makedirs(logdir, exist_ok=True)
    with open(os.path.join(logdir, "ptxas.log"), "w") as fh:
        fh.write(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
