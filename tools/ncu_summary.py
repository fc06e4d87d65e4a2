#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep (raw page) into the few lines the design cares about.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [substring ...]"""
import csv
import subprocess
import sys

DEFAULT = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active", "sm__inst_executed_pipe_fp64", "sm__inst_executed_pipe_lsu",
    "sm__inst_executed_pipe_alu.avg.pct", "sm__inst_executed_pipe_xu.avg.pct", "sm__inst_executed_pipe_fma.avg.pct",
    "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
    "smsp__average_warp", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed",
]


def main():
    rep = sys.argv[1]
    want = sys.argv[2:] or DEFAULT
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print("== kernel:", name)
        for h, u, v in zip(hdr, units, vals):
            if any(w in h for w in want) and "realtime" not in h:
                print(f"{h:95s} {u:14s} {v}")


if __name__ == "__main__":
    main()
