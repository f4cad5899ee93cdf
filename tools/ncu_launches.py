#!/usr/bin/env python
"""Summarise an `ncu --csv --metrics ...` launch log: one line per kernel launch."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hdr]
ki, gi, mi, vi = h.index("Kernel Name"), h.index("Grid Size"), h.index("Metric Name"), h.index("Metric Value")
cur = {}
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    cur.setdefault((int(r[0]), r[ki][:52], r[gi]), {})[r[mi]] = r[vi].replace(",", "")
tot = 0.0
for (idx, name, grid), v in sorted(cur.items()):
    t = float(v.get("gpu__time_duration.sum", 0))
    tot += t
    extra = "  ".join(f"{k.split('.')[0][-28:]}={float(x):.1f}" for k, x in v.items() if k != "gpu__time_duration.sum")
    print(f"{idx:4d} {name:52s} {grid:14s} {t / 1e3:8.2f} us  {extra}")
print(f"total {tot / 1e3:.1f} us over {len(cur)} launches")
