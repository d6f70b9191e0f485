"""TEST INFRASTRUCTURE ONLY.  Compiles oracle/rvq_oracle.c into oracle/_build/librvq_oracle.so.

Called by __graft_entry__.build() (building the checker is not using it) and lazily by
oracle/c_oracle.py when the shared object is missing or older than the source.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "rvq_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "librvq_oracle.so")

# -ffp-contract=off: every rounding in the restatement is explicit (fmaf where the reference fuses).
# -mfma -mavx2: fmaf() becomes one instruction; any x86-64-v3 host can run the result.
CFLAGS = ["-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-mfma", "-mavx2", "-fPIC", "-shared", "-Wall"]


def needs_build() -> bool:
    return (not os.path.exists(OUT)) or os.path.getmtime(OUT) < os.path.getmtime(SRC)


def build(force: bool = False) -> str:
    if force or needs_build():
        os.makedirs(OUT_DIR, exist_ok=True)
        cmd = ["gcc", *CFLAGS, SRC, "-o", OUT, "-lm"]
        subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
