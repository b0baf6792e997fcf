#!/usr/bin/env python3
"""Hot-loop view of an ncu source-page CSV of a token-parse kernel: per segment of equally often executed
instructions, the instruction count and the cycles it costs per loop trip; then the stall reasons per trip.
usage: ncu_hot.py source.csv cycles_per_trip [min_exec_fraction]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1]))); trip = float(sys.argv[2]); thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.02
h = rows[1]; data = rows[2:]
isrc, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
tot = sum(int(r[isamp]) for r in data); mx = max(int(r[iex]) for r in data)
unit = tot / trip
seg = []; cur = None
for k, r in enumerate(data):
    ex = int(r[iex]) / mx
    if ex < thr: cur = None; continue
    key = round(ex, 2)
    if cur is None or abs(cur[4] - key) > 0.03: cur = [k, k, 0, 0, key]; seg.append(cur)
    cur[1] = k; cur[2] += int(r[isamp]); cur[3] += 1
cold = tot - sum(s[2] for s in seg)
for s in seg: print(f"{s[0]:5d}-{s[1]:5d} exec {s[4]:.2f} instrs {s[3]:3d} cycles {s[2]/unit:7.1f}")
print(f"below threshold: cycles {cold/unit:.1f}")
stalls = [(i, c) for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
acc = {}
for r in data:
    for i, c in stalls: acc[c] = acc.get(c, 0) + int(r[i])
print({c[6:]: round(v / unit, 1) for c, v in sorted(acc.items(), key=lambda x: -x[1])[:8]})
if len(sys.argv) > 4:
    lo, hi = int(sys.argv[4]), int(sys.argv[5])
    for k in range(lo, hi):
        r = data[k]
        st = sorted(((int(r[i]), c) for i, c in stalls), reverse=True)[:2]
        print(f"{k:5d} {int(r[iex])/mx:5.2f} {int(r[isamp])/unit:6.1f} {r[isrc].strip()[:64]:64s} {st[0][1][6:]}:{st[0][0]/unit:.1f} {st[1][1][6:]}:{st[1][0]/unit:.1f}")
