#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libvrvq.so (cuobjdump -sass; runs without a GPU).

  python scripts/sass_opcodes.py [> profiles/rN_sass_opcodes.txt]

Counts the mnemonics that prove which hardware path a kernel uses: UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st),
UTCBAR (tcgen05.commit), UBLKCP (cp.async.bulk), UTMALDG (cp.async.bulk.tensor = TMA tiled loads), SYNCS (mbarrier), and the
CUDA-core FP32 work (FFMA, FFMA2, FMNMX*), plus the total instruction count.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vrvq_b200", "libvrvq.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMAPF", "SYNCS", "LDG", "STG", "LDS", "STS", "LDGSTS",
        "FFMA2", "FFMA", "FMNMX3", "FMNMX", "HMMA", "BAR", "MUFU"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            kernels[cur][op] += 1
    demangled = {}
    try:
        names = list(kernels)
        d = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
        demangled = dict(zip(names, d))
    except Exception:
        pass
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: opcode counts per kernel (static instruction counts)")
    for k, c in kernels.items():
        name = demangled.get(k, k)
        name = name.rsplit(">(", 1)[0] + ">" if ">(" in name else name.split("(")[0]
        name = name.replace("(int)", "").replace("(bool)1", "true").replace("(bool)0", "false")
        cols = [f"{key}={sum(v for op, v in c.items() if op == key or (key in ('LDG', 'STG', 'LDS', 'STS') and op.startswith(key) and not op.startswith('LDGSTS')))}" for key in KEYS]
        cols = [x for x in cols if not x.endswith("=0")]
        print(f"{name}\n    total={c['_total']} " + " ".join(cols))


if __name__ == "__main__":
    sys.exit(main())
