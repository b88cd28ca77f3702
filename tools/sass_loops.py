#!/usr/bin/env python
"""Lists the loops (backward branches) of one kernel in a `cuobjdump -sass` dump with their instruction mix.
usage: sass_loops.py dump.sass mangled_kernel_name [min_instructions]"""
import collections
import re
import sys

path, name = sys.argv[1], sys.argv[2]
minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ins = []
on = False
for line in open(path):
    if "Function :" in line:
        on = line.strip().endswith(name)
        continue
    if not on:
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
print("%s: %d instructions" % (name, len(ins)))
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", t)
    if not m:
        continue
    tgt = int(m.group(1), 16)
    if tgt < a and tgt in addr and i - addr[tgt] >= minlen:
        body = ins[addr[tgt]:i + 1]
        h = collections.Counter()
        for _, x in body:
            x = re.sub(r"^@!?U?P\d+\s+", "", x)
            op = x.split()[0]
            h[op.split(".")[0]] += 1
        if len(body) < 400: print("loop 0x%x..0x%x: %d instr: %s" % (tgt, a, len(body), ", ".join("%s %d" % kv for kv in h.most_common(16))))
