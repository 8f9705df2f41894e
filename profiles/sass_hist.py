#!/usr/bin/env python3
"""Summarise `ncu -i X.ncu-rep --page source --csv` (SASS view): executed warp-instructions per opcode and the
hottest stall-sample lines of the FIRST kernel instance in the report."""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; data = []; seen = 0
for r in rows:
    if r and r[0] == "Address":
        seen += 1
        hdr = r
        continue
    if seen == 1 and hdr and len(r) == len(hdr):
        data.append(r)
iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[iE]) for r in data); totS = sum(int(r[iN]) for r in data)
print("warp instructions executed:", tot, " stall samples:", totS, " SASS lines:", len(data))
op = defaultdict(int); ops = defaultdict(int)
for r in data:
    t = r[iS].split()
    o = t[1] if t[0].startswith("@") else t[0]
    op[o] += int(r[iE]); ops[o] += int(r[iN])
for k, v in sorted(op.items(), key=lambda x: -x[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 24]:
    print(f"{k:28s} {v:12d} {100*v/tot:5.1f}%   samples {100*ops[k]/max(totS,1):5.1f}%")
print("-- hottest lines by samples")
for r in sorted(data, key=lambda r: -int(r[iN]))[:12]:
    print(f"{int(r[iN]):7d} {int(r[iE]):10d}  {r[iS].strip()}")
