#!/usr/bin/env python3
"""Summarise .ncu-rep files (read with `ncu -i ... --page raw --csv`, no GPU needed) into a markdown table."""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
    ("smsp__sass_average_branch_targets_threads_uniform.pct", "branch uniformity %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 sector hit rate %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 sector hit rate %"),
    ("lts__t_sectors.avg.pct_of_peak_sustained_elapsed", "L2 sector throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard (warps/issue)"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
]


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [(dict(zip(hdr, r)), dict(zip(hdr, units))) for r in rows[2:]]


def main():
    for path in sys.argv[1:]:
        for vals, units in load(path):
            print(f"\n### `{vals.get('Kernel Name', '?')}`  ({path.split('/')[-1]})\n")
            print("| metric | value |\n|---|---|")
            for key, label in METRICS:
                if key in vals:
                    print(f"| {label} (`{key}`) | {vals[key]} {units.get(key, '')} |")


if __name__ == "__main__":
    main()
