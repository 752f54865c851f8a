#!/usr/bin/env python
"""Extract the per-launch DRAM traffic (and a few deciding metrics) of the first kernel in an `ncu --page raw --csv`
dump into a small JSON that bench.py reads for roofline.traffic.
    python tools/ncu_traffic.py profiles/r02_head_conv_ncu_full_raw.csv profiles/r02_head_conv_traffic.json
"""
import csv
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(src, dst):
    rows = list(csv.reader(open(src)))
    head, units, vals = rows[0], rows[1], rows[2]
    col = {n: i for i, n in enumerate(head)}

    def get(name):
        i = col[name]
        return float(vals[i].replace(",", "")) * UNIT.get(units[i], 1.0)

    out = {
        "kernel": vals[col["Kernel Name"]],
        "dram_bytes_read": get("dram__bytes_read.sum"),
        "dram_bytes_write": get("dram__bytes_write.sum"),
        "gpu_time_ms_under_ncu": float(vals[col["gpu__time_duration.sum"]]) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(units[col["gpu__time_duration.sum"]], 1.0),
        "source": f"ncu --set full --clock-control none, {src} (dram__bytes_read.sum + dram__bytes_write.sum of one launch)",
    }
    for k in ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
              "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed"):
        if k in col:
            try:
                out[k.split(".TriageCompute.")[-1]] = float(vals[col[k]])
            except ValueError:
                pass
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
