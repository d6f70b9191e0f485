"""Summarise `ncu --page source --csv` output: instruction mix by opcode, stall reasons, hottest SASS lines."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
data = []
nk = 0
for r in rows:
    if len(r) > 5 and r[0] == "Address":
        hdr = r
        nk += 1
        continue
    if hdr and len(r) == len(hdr) and nk == 1:
        data.append(r)
iS, iE, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
iWf, iWfI = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ops, samp, wf = collections.Counter(), collections.Counter(), collections.Counter()
tot = 0
for r in data:
    src = r[iS].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    full = m.group(2) if m else src[:12]
    op = ".".join(full.split(".")[:2]) if full.startswith(("LDS", "STS", "LDG", "STG")) else full.split(".")[0]
    n = int(r[iE])
    ops[op] += n
    tot += n
    samp[op] += int(r[iSamp])
    wf[op] += int(r[iWf])
print("total warp-instructions", tot, " total samples", sum(samp.values()))
for op, n in ops.most_common(28):
    print(f"{op:14s} {n:12d} {100*n/tot:5.1f}%  samples {samp[op]:7d}  smem wavefronts {wf[op]}")
st = collections.Counter()
for r in data:
    for i in stall_cols:
        st[hdr[i]] += int(r[i])
print("stalls:", [(k, v) for k, v in st.most_common(10)])
top = sorted(data, key=lambda r: -int(r[iSamp]))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]
print("hottest SASS lines (samples, executed, text):")
for r in top:
    print(f"  {int(r[iSamp]):6d} {int(r[iE]):10d}  {r[iS].strip()[:90]}")
