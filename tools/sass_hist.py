"""Summarise an `ncu --page source --csv` dump: executed-instruction histogram and the hot SASS ranges."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iA, iE, iS = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = rows[2:]
tot = sum(int(r[iE]) for r in data)
tots = sum(int(r[iS]) for r in data)
print("total warp instructions", tot, "samples", tots)
c = Counter()
for r in data:
    c[round(int(r[iE]) / 1e6)] += 1
for k, v in sorted(c.items()):
    print(f"{k:8d} M x {v:4d} instrs = {k * v / 1e3:8.2f} G  ({100.0 * k * v * 1e6 / tot:5.1f} %)")
if len(sys.argv) > 2:
    lo, hi = float(sys.argv[2]) * 1e6, float(sys.argv[3]) * 1e6
    for n, r in enumerate(data):
        e = int(r[iE])
        if lo <= e <= hi:
            print(n, r[iA][:90], e // 1000000, r[iS])
