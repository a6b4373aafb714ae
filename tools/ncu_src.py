#!/usr/bin/env python
"""Per-instruction stall summary from `ncu -i X.ncu-rep --page source --csv --kernel-name regex:K`.
  python tools/ncu_src.py file.csv [min_exec]   -> offset, executed, samples, top stalls, SASS"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
nxt = next((i for i, r in enumerate(rows) if i > hdr_i and r and r[0] == "Kernel Name"), len(rows))
rows = rows[:nxt]  # first kernel instance only
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
minexec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
base = None
tot = 0
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr): continue
    tot += int(r[col["# Samples"]] or 0)
print("total samples", tot)
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr): continue
    a = int(r[0], 16)
    if base is None: base = a
    ex = int(r[col["Instructions Executed"]] or 0)
    if ex < minexec: continue
    smp = int(r[col["# Samples"]] or 0)
    st = sorted(((int(r[col[s]] or 0), s[6:]) for s in stalls), reverse=True)[:3]
    print(f"{a-base:5x} {ex:9d} {smp:6d} {100*smp/tot:5.1f}% " + " ".join(f"{n}:{v}" for v, n in st if v) + "  | " + r[1].strip())
