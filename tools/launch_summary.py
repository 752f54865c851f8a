#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count and total time.
    python tools/launch_summary.py gpurun_out/launches.csv [--md]"""
import collections
import csv
import re
import sys


def load(path):
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3}[u]
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("ss::", "")
        name = re.sub(r"at::native::|<unnamed>::", "", name)
        rows.append((name[:70], ms, r.get("Grid Size", ""), r.get("Block Size", "")))
    return rows


def main():
    path = sys.argv[1]
    rows = load(path)
    agg = collections.OrderedDict()
    for n, ms, *_ in rows:
        c, t = agg.get(n, (0, 0.0))
        agg[n] = (c + 1, t + ms)
    tot = sum(t for _, t in agg.values())
    print(f"{len(rows)} launches, {tot:.3f} ms serialised (cold-cache, per-launch ncu replay)")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {t:8.3f} ms  {100*t/tot:5.1f} %  x{c:<4d} {n}")


if __name__ == "__main__":
    main()
