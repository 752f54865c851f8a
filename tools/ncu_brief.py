#!/usr/bin/env python
"""Print the handful of ncu --page details metrics that decide what bounds a kernel."""
import csv, sys
KEYS = ["Duration", "DRAM Throughput", "Memory Throughput", "Compute (SM) Throughput", "L2 Cache Throughput", "L1/TEX Cache Throughput",
        "Executed Ipc Active", "Issue Slots Busy", "Registers Per Thread", "Dynamic Shared Memory Per Block", "Achieved Occupancy",
        "Theoretical Occupancy", "Mem Busy", "Max Bandwidth", "Mem Pipes Busy", "One or More Eligible", "No Eligible",
        "L2 Hit Rate", "Warp Cycles Per Issued Instruction", "Block Limit Shared Mem", "Waves Per SM", "SM Frequency", "DRAM Frequency"]
for path in sys.argv[1:]:
    rows = list(csv.DictReader(open(path)))
    kern = {}
    for r in rows:
        kern.setdefault((r["ID"], r["Kernel Name"][:50]), []).append(r)
    for (kid, name), rs in kern.items():
        print(f"== {path} :: {name}")
        for r in rs:
            if r["Metric Name"] in KEYS:
                print(f"   {r['Section Name'][:28]:28s} {r['Metric Name']:38s} {r['Metric Value']:>14s} {r['Metric Unit']}")
