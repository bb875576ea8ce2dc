#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep (raw page) into the handful of counters the roofline discussion needs.
usage: ncu_summary.py <file.ncu-rep> [launch index]"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sectors.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active"]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, row = rows[0], rows[1], rows[2 + which]
    print("kernel:", row[hdr.index("Kernel Name")][:100])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:82s} {row[i]:>16s} {units[i]}")
    stalls = [(float(row[i].replace(",", "")), h[len(STALL):-len("_per_issue_active.ratio")]) for i, h in enumerate(hdr)
              if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and row[i]]
    print("warp stall reasons (warps per issue-active cycle):")
    for v, n in sorted(stalls, reverse=True)[:8]:
        print(f"    {n:40s} {v:8.2f}")


if __name__ == "__main__":
    main()
