#!/usr/bin/env python
"""Condenses an .ncu-rep (ncu --set full) into the CSV kept under profiles/: one row per captured launch, the metrics the
design discussion uses.  usage: tools/ncu_export.py gpurun_out/x.ncu-rep profiles/x.csv"""
import csv
import subprocess
import sys

KEEP = ["ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]


def main():
    rep, out = sys.argv[1:3]
    text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(text.splitlines()))
    hdr = rows[0]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        for r in rows:
            if len(r) >= len(hdr):
                w.writerow([r[i] for i in idx])
    print("wrote", out, len(rows) - 2, "launches")


if __name__ == "__main__":
    main()
