#!/usr/bin/env python3
"""Share of device time per kernel from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value")
t = defaultdict(float); n = defaultdict(int)
for r in rows[1:]:
    name = r[ki].split("(")[0]
    t[name] += float(r[vi].replace(",", "")); n[name] += 1
tot = sum(t.values())
print(f"{len(rows)-1} launches, {tot/1e6:.2f} ms of kernel time (cold-cache, serialised: compare SHARES)")
for k, v in sorted(t.items(), key=lambda kv: -kv[1]):
    print(f"{v/1e6:10.2f} ms  {100*v/tot:5.1f}%  x{n[k]:3d}  {k}")
