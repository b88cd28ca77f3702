#!/usr/bin/env python
"""Condenses `ncu --page raw --csv` exports into the files bench.py and DESIGN.md cite.

usage: ncu_summary.py <workload> <problems> <raw.csv> [<workload> <problems> <raw.csv> ...]
  profiles/r2_<workload>_ncu.csv   one row per launch of the step: the metrics that matter for this kernel
  profiles/r2_traffic.json         DRAM bytes per step per workload (roofline.traffic in bench.py)
The raw CSV holds the solve-kernel launches of `bench.py --kernel-only --workload W` (warm-up passes + one step); the
rows of the last pass are kept.
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    "Kernel Name", "launch__grid_size", "launch__registers_per_thread", "gpu__time_duration.sum", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def last_step(hdr, data):
    """The capture holds warm-up passes and the timed step: every pass launches the same sequence of (kernel, grid),
    so the rows of the last pass are the last p rows, p the smallest period of that sequence."""
    key = [(r[hdr.index("Kernel Name")], r[hdr.index("launch__grid_size")]) for r in data]
    for p in range(1, len(key) + 1):
        if len(key) % p == 0 and all(key[i] == key[i % p] for i in range(len(key))):
            return data[-p:]
    return data


def main():
    out_json = os.path.join(ROOT, "profiles", "r2_traffic.json")
    traffic = json.load(open(out_json)) if os.path.exists(out_json) else {}
    args = sys.argv[1:]
    for k in range(0, len(args), 3):
        workload, problems, path = args[k], int(args[k + 1]), args[k + 2]
        rows = list(csv.reader(open(path)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        data = last_step(hdr, data)
        cols = [hdr.index(c) for c in KEEP if c in hdr]
        dst = os.path.join(ROOT, "profiles", "r2_%s_ncu.csv" % workload)
        with open(dst, "w", newline="") as f:
            wr = csv.writer(f)
            wr.writerow([hdr[c] for c in cols])
            wr.writerow([units[c] for c in cols])
            for r in data:
                wr.writerow([r[c] for c in cols])
        total = 0.0
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            c = hdr.index(name)
            total += sum(float(r[c].replace(",", "")) * SCALE[units[c]] for r in data)
        traffic[workload] = {"problems": problems, "dram_bytes_per_step": total, "launches": len(data),
                             "source": "profiles/%s (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum over the step's %d launches)"
                                       % (os.path.basename(dst), len(data))}
        print(workload, traffic[workload])
    json.dump(traffic, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main()
