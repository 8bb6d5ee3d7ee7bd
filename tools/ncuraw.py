#!/usr/bin/env python
"""Compact per-kernel table from `ncu -i rep --page raw --csv`."""
import csv
import subprocess
import sys


def main(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, data = rows[0], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def g(r, k):
        try:
            return float(r[col[k]].replace(",", ""))
        except Exception:
            return float("nan")

    print("%-40s %-13s %-6s %7s %7s %6s %6s %6s %4s %6s %6s %6s %6s %8s %7s %7s %5s" % (
        "kernel", "grid", "block", "us", "Minst", "issue%", "dram%", "warps%", "regs", "longSB", "shrtSB", "barr", "mio",
        "bankconf", "rdMB", "wrMB", "L1hit"))
    for r in data:
        print("%-40s %-13s %-6s %7.1f %7.2f %6.1f %6.1f %6.1f %4d %6.2f %6.2f %6.2f %6.2f %8d %7.1f %7.1f %5.0f" % (
            r[col["Kernel Name"]][:40], r[col["Grid Size"]].replace(" ", ""), r[col["Block Size"]].split(",")[0].strip("("),
            g(r, "gpu__time_duration.sum"), g(r, "smsp__inst_executed.sum") / 1e6,
            g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            g(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), g(r, "launch__registers_per_thread"),
            g(r, "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
            g(r, "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
            g(r, "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
            g(r, "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
            g(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"), g(r, "dram__bytes_read.sum"),
            g(r, "dram__bytes_write.sum"), g(r, "l1tex__t_sector_hit_rate.pct")))


if __name__ == "__main__":
    main(sys.argv[1])
