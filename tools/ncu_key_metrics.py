#!/usr/bin/env python3
"""Key metrics of an ncu --set full report as text (one block per profiled launch), for profiles/.
Usage: ncu_key_metrics.py <report.ncu-rep> [header line]"""
import csv
import io
import subprocess
import sys

KEEP = ["Duration", "Elapsed Cycles", "SM Frequency", "DRAM Throughput", "Memory Throughput", "L1/TEX Cache Throughput",
        "L2 Cache Throughput", "Compute (SM) Throughput", "Executed Ipc Active", "Issue Slots Busy", "SM Busy", "L1/TEX Hit Rate",
        "L2 Hit Rate", "Warp Cycles Per Issued Instruction", "Avg. Active Threads Per Warp", "Executed Instructions",
        "Block Size", "Grid Size", "Registers Per Thread", "Dynamic Shared Memory Per Block", "Static Shared Memory Per Block",
        "Block Limit Registers", "Block Limit Shared Mem", "Theoretical Active Warps per SM", "Achieved Active Warps Per SM",
        "Theoretical Occupancy", "Achieved Occupancy"]


def main():
    rep = sys.argv[1]
    if len(sys.argv) > 2:
        print(sys.argv[2])
    out = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.DictReader(io.StringIO(out)))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics",
                          "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed_pipe_fp64.sum,"
                          "sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"],
                         capture_output=True, text=True).stdout
    rawrows = list(csv.DictReader(io.StringIO(raw)))
    by_id = {}
    for r in rows:
        by_id.setdefault(r["ID"], []).append(r)
    for kid, rs in by_id.items():
        print(f"\n== {rs[0]['Kernel Name']}  grid {rs[0].get('Grid Size', '')} block {rs[0].get('Block Size', '')}")
        seen = set()
        for r in rs:
            name = r["Metric Name"]
            if name in KEEP and (name, r["Section Name"]) not in seen:
                seen.add((name, r["Section Name"]))
                print(f"   {name:45s} {r['Metric Value']:>16s} {r['Metric Unit']}")
        for rr in rawrows:
            if rr.get("ID") == kid:
                for k, v in rr.items():
                    if k.startswith("dram__bytes") or k.startswith("sm__pipe_fp64") or k.startswith("sm__inst_executed_pipe_fp64"):
                        print(f"   {k:45s} {v:>16s}")


if __name__ == "__main__":
    main()
