#!/usr/bin/env python
"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel share table."""
import collections, csv, re, sys

def main(path, out):
    lines = open(path).readlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0; big = []
    for r in csv.DictReader(lines[start:]):
        v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
        v = v / 1e6 if u == "ns" else v / 1e3 if u == "us" else v
        short = re.sub(r"\(.*", "", r["Kernel Name"])[:90]
        agg[short][0] += 1; agg[short][1] += v; tot += v
        if v > 0.3: big.append((v, r.get("Grid Size", ""), short))
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary: {path}\n\ntotal {tot:.3f} ms over {sum(n for n, _ in agg.values())} launches "
                "(one forward step, config2; cold-cache serialised launches: compare SHARES)\n\n| ms | share | launches | kernel |\n|---|---|---|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
            f.write(f"| {t:.3f} | {100 * t / tot:.1f}% | {n} | `{k}` |\n")
        f.write("\nLaunches over 0.3 ms:\n\n| ms | grid | kernel |\n|---|---|---|\n")
        for v, g, k in big:
            f.write(f"| {v:.3f} | {g} | `{k}` |\n")

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
