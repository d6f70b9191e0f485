"""Join `ncu --page source --csv` (SASS view) with `nvdisasm -g -c` line info: per source line, instructions executed,
stall samples and shared-memory wavefronts.  usage: ncu_by_line.py <ncu_sass.csv> <nvdisasm.sass> <mangled-kernel-substring> [top]"""
import collections
import csv
import re
import sys

ncu_csv, sass_file, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# ---- nvdisasm: list of (line_no) per instruction, in order, for the kernel
lines = open(sass_file).read().split("\n")
in_k = False
cur_line = None
seq = []  # (opcode_text, line)
for ln in lines:
    if ln.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", ln):
        in_k = kname in ln
        continue
    if not in_k:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        seq.append((m.group(2).strip(), cur_line))
# ---- ncu rows
rows = list(csv.reader(open(ncu_csv)))
hdr, data, nk = None, [], 0
for r in rows:
    if len(r) > 5 and r[0] == "Address":
        hdr = r
        nk += 1
        continue
    if hdr and len(r) == len(hdr) and nk == 1:
        data.append(r)
iS, iE, iSamp, iWf = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("L1 Wavefronts Shared")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
print(f"nvdisasm instructions: {len(seq)}  ncu instructions: {len(data)}")
n = min(len(seq), len(data))
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
for k in range(n):
    key = seq[k][1]
    a = agg[key]
    a[0] += int(data[k][iE])
    a[1] += int(data[k][iSamp])
    a[2] += int(data[k][iWf])
    for i, h in stall_cols:
        v = int(data[k][i])
        if v:
            a[3][h[6:]] += v
totE = sum(a[0] for a in agg.values())
totS = sum(a[1] for a in agg.values())
src_cache = {}
def src(key):
    if key is None:
        return ""
    f, l = key
    if f not in src_cache:
        try:
            src_cache[f] = open(f"/root/repo/vrvq_b200/csrc/{f}").read().split("\n")
        except Exception:
            src_cache[f] = []
    t = src_cache[f]
    return t[l - 1].strip()[:70] if 0 < l <= len(t) else ""
print(f"{'line':>18s} {'instr%':>7s} {'samp%':>6s} {'wavefr':>9s}  top stalls | source")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    st = ",".join(f"{k}:{v}" for k, v in a[3].most_common(3))
    name = f"{key[0]}:{key[1]}" if key else "?"
    print(f"{name:>18s} {100*a[0]/totE:6.1f}% {100*a[1]/totS:5.1f}% {a[2]:9d}  {st:40s} | {src(key)}")

# ---- optional: aggregate by named line ranges given as extra args  name:lo-hi
if len(sys.argv) > 5:
    print("\nby region (rvq_encode.cu line ranges):")
    regs = []
    for spec in sys.argv[5:]:
        nm, rng = spec.split(":")
        lo, hi = map(int, rng.split("-"))
        regs.append((nm, lo, hi))
    out = collections.OrderedDict((nm, [0, 0, 0, collections.Counter()]) for nm, _, _ in regs)
    out["other"] = [0, 0, 0, collections.Counter()]
    for key, a in agg.items():
        tgt = "other"
        if key and key[0] == "rvq_encode.cu":
            for nm, lo, hi in regs:
                if lo <= key[1] <= hi:
                    tgt = nm
                    break
        o = out[tgt]
        o[0] += a[0]; o[1] += a[1]; o[2] += a[2]; o[3].update(a[3])
    for nm, o in out.items():
        st = ",".join(f"{k}:{v}" for k, v in o[3].most_common(4))
        print(f"{nm:12s} instr {100*o[0]/totE:5.1f}%  samples {100*o[1]/totS:5.1f}%  smem wavefronts {o[2]:10d}  {st}")
