#!/usr/bin/env python
"""Text summary of an `ncu --set full` report: per launch the headline metrics of the details page plus the raw
DRAM / L2 / instruction counters.  usage: python profiles/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_x.txt"""
import collections
import csv
import io
import subprocess
import sys

WANT = ["Duration", "DRAM Throughput", "Memory Throughput", "L2 Cache Throughput", "Executed Ipc Active", "Issue Slots Busy",
        "Registers Per Thread", "Achieved Occupancy", "Theoretical Occupancy", "Avg. Active Threads Per Warp",
        "Avg. Not Predicated Off Threads Per Warp", "Executed Instructions", "No Eligible", "L1/TEX Hit Rate", "L2 Hit Rate",
        "Mem Busy", "Max Bandwidth", "Grid Size", "Block Size", "Dynamic Shared Memory Per Block", "Static Shared Memory Per Block"]
RAW = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "lts__t_sectors.sum", "smsp__inst_executed.sum",
       "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def page(rep, name):
    return subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout


def main(rep):
    rows = list(csv.DictReader(io.StringIO(page(rep, "details"))))
    by = collections.OrderedDict()
    for r in rows:
        by.setdefault((r["ID"], r["Kernel Name"]), {})[r["Metric Name"]] = (r["Metric Value"], r["Metric Unit"])
    raw = list(csv.reader(io.StringIO(page(rep, "raw"))))
    h, units = raw[0], raw[1]
    col = {n: i for i, n in enumerate(h)}
    print("ncu --set full summary of %s" % rep)
    for n, ((lid, kname), m) in enumerate(by.items()):
        print("\nlaunch %s  %s" % (lid, kname[:100]))
        for w in WANT:
            if w in m:
                print("  %-44s %s %s" % (w, m[w][0], m[w][1]))
        r = raw[2 + n]
        for w in RAW:
            if w in col:
                print("  %-44s %s %s" % (w, r[col[w]], units[col[w]]))


if __name__ == "__main__":
    main(sys.argv[1])
