#!/usr/bin/env python3
"""Selected metrics of every kernel in an ncu report (raw page).  usage: ncu_summary.py report.ncu-rep"""
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
           "launch__shared_mem_per_block_static", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed_op_local_st.sum",
           "smsp__inst_executed_op_local_ld.sum"]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    name = hdr.index("Kernel Name")
    for r in rows[2:]:
        print(f"Kernel Name = {r[name]} ")
        for m in METRICS:
            if m in hdr:
                k = hdr.index(m)
                print(f"{m} = {r[k]} {units[k]}")
        print()


if __name__ == "__main__":
    main()
