#!/usr/bin/env python
"""Top source lines of an `ncu -i x.ncu-rep --page source --csv` dump by stall samples.
    python tools/ncu_source_top.py dump.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
for i, r in enumerate(rows):
    if any("Samples" in c for c in r):
        hdr = i
        break
if hdr is None:
    print("no header found"); sys.exit(1)
h = rows[hdr]
col = next(i for i, c in enumerate(h) if "Warp Stall Sampling (All" in c or c.strip() == "# Samples" or "Sampling (All" in c)
src = next((i for i, c in enumerate(h) if c.strip() in ("Source", "SASS")), 1)
data = []
for r in rows[hdr + 1:]:
    try:
        data.append((float(r[col].replace(",", "") or 0), r))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data) or 1.0
print("columns:", [c for c in h][:12])
for v, r in sorted(data, key=lambda t: -t[0])[:n]:
    print(f"{v:8.0f} {100*v/tot:5.1f}%  {r[0][:6]:>6s}  {r[src][:150]}")
