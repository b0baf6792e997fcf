#!/usr/bin/env python3
"""Per-instruction view of an ncu source page CSV: executed count (in units of `unit`), samples, top stall."""
import csv, sys
path = sys.argv[1]; unit = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0; hi = int(sys.argv[4]) if len(sys.argv) > 4 else 10**9
rows = list(csv.reader(open(path)))
h = rows[1]; data = rows[2:]
ia, isrc, isamp, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stalls = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r[iex]) for r in data); tots = sum(int(r[isamp]) for r in data)
print(f"total executed {tot/unit:.1f} units, samples {tots}")
for k, r in enumerate(data):
    if not (lo <= k < hi): continue
    st = sorted(((int(r[i]), c) for i, c in stalls), reverse=True)[:2]
    print(f"{k:4d} {int(r[iex])/unit:8.2f} {int(r[isamp]):7d} {r[isrc].strip()[:70]:70s} {st[0][1][6:]}:{st[0][0]} {st[1][1][6:]}:{st[1][0]}")
