#!/usr/bin/env python
"""Condenses an `ncu --page source --csv` export (SASS view, one row per instruction with warp-stall samples) into one
line per loop: share of the kernel's samples, instructions executed, top stall reasons.
usage: ncu_source_summary.py src.csv [min_loop_instructions]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
minlen = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print("#", rows[0][0], rows[0][1])
h, data = rows[1], rows[2:]
ia, isrc, ismp, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
addr = [int(r[ia], 16) if r[ia].lower().startswith("0x") else int(r[ia]) for r in data]
base = addr[0]
pos = {a: i for i, a in enumerate(addr)}
smp = [int(r[ismp] or 0) for r in data]
total = sum(smp)
print("# %d instructions, %d stall samples" % (len(data), total))
loops = []
for i, r in enumerate(data):
    m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", r[isrc])
    if not m:
        continue
    tgt = int(m.group(1), 16) + (base if int(m.group(1), 16) + base in pos else 0)
    if tgt in pos and pos[tgt] < i and i - pos[tgt] + 1 >= minlen:
        loops.append((pos[tgt], i))
# innermost loops only
inner = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
covered = 0
for lo, hi in inner:
    s = sum(smp[lo:hi + 1])
    covered += s
    if s < 0.005 * total:
        continue
    ex = max(int(data[k][iex] or 0) for k in range(lo, hi + 1))
    st = collections.Counter()
    for k in range(lo, hi + 1):
        for i, c in stall:
            v = data[k][i]
            if v:
                st[c] += int(v)
    top = ", ".join("%s %.0f%%" % (c.replace("stall_", ""), 100.0 * v / max(1, sum(st.values()))) for c, v in st.most_common(4))
    mix = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", data[k][isrc]).split()[0].split(".")[0] for k in range(lo, hi + 1))
    print("loop +0x%x..+0x%x  %3d instr  %5.1f%% of samples  trips(warp) %d  stalls: %s  | %s"
          % (addr[lo] - base, addr[hi] - base, hi - lo + 1, 100.0 * s / total, ex, top, ", ".join("%s %d" % kv for kv in mix.most_common(8))))
print("# outside these loops: %.1f%% of samples" % (100.0 * (total - covered) / total))
