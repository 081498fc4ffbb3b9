#!/usr/bin/env python
"""Key metrics of `.ncu-rep` captures as one CSV (the form committed under profiles/).

    python tools/ncu_extract.py out.csv label=path.ncu-rep [label=path.ncu-rep ...]

Reads each report with `ncu -i <rep> --page raw --csv` (works on the CPU container) and keeps the columns below.
"""
import csv
import io
import subprocess
import sys

KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__shared_mem_per_block_dynamic",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def main():
    out, pairs = sys.argv[1], [a.rsplit("=", 1) for a in sys.argv[2:]]
    rows, units = [], None
    for label, path in pairs:
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rd = list(csv.reader(io.StringIO(txt)))
        head = rd[0]
        idx = [head.index(k) if k in head else -1 for k in KEEP]
        if units is None:
            units = [""] + [rd[1][i] if i >= 0 else "" for i in idx]
        for r in rd[2:]:
            rows.append([label] + [r[i] if i >= 0 else "" for i in idx])
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["case"] + KEEP)
        w.writerow(units)
        w.writerows(rows)
    print(f"{out}: {len(rows)} launches")


if __name__ == "__main__":
    main()
