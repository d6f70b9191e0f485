"""Builds vrvq_b200/libvrvq.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

python -m vrvq_b200.build [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libvrvq.so")
SOURCES = ["abi.cu", "rvq_encode.cu", "rvq_encode_tc.cu", "rvq_aux.cu", "wire.cu", "subnet.cu", "subnet_tc.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "encode_params.cuh"), os.path.join(CSRC, "tmaps.cuh"), os.path.join(os.path.dirname(HERE), "include", "vrvq.h")]
OBJ_DIR = os.path.join(CSRC, "_obj")  # per-source objects (git-ignored): only changed sources are recompiled

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false",  # nothing is contracted behind our back; every FMA in the kernels is an explicit fmaf
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2",
]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build() -> bool:
    return _stale(OUT, [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)])


def build(force: bool = False, verbose: bool = False) -> str:
    if not (force or needs_build()):
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    common = HEADERS + [os.path.abspath(__file__)]
    jobs = []
    objs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ_DIR, os.path.splitext(s)[0] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + common):
            cmd = [nvcc, *NVCC_FLAGS] + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))  # all sources in parallel
    failed = []
    for s, proc in jobs:
        out, _ = proc.communicate()
        if verbose or proc.returncode != 0:
            sys.stderr.write(out)
        if proc.returncode != 0:
            failed.append(s)
    if failed:
        raise RuntimeError(f"nvcc failed compiling {', '.join(failed)}")
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *objs, "-o", OUT], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed linking libvrvq.so")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
