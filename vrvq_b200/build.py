"""Builds vrvq_b200/libvrvq.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

python -m vrvq_b200.build [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libvrvq.so")
SOURCES = ["abi.cu", "rvq_encode.cu", "rvq_encode_tc.cu", "rvq_aux.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "encode_params.cuh"), os.path.join(os.path.dirname(HERE), "include", "vrvq.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false",  # nothing is contracted behind our back; every FMA in the kernels is an explicit fmaf
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2",
    "-shared",
]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not (force or needs_build()):
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-o", OUT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libvrvq.so")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
