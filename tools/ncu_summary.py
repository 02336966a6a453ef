#!/usr/bin/env python3
"""Turns ncu output into the markdown summaries kept under profiles/.

    python tools/ncu_summary.py launches <launches.csv> <out.md> "<command>" [note]
    python tools/ncu_summary.py full <report.ncu-rep> <out.md> "<command>" [note]
"""
import collections
import csv
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def launches(path, out, command, note):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if len(r) > 5]
    h = rows[0]
    ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
        a = agg.setdefault(r[ik], [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list\n\nCommand: `{command}`\n{note}\n\n| kernel | launches | total ms | share of all device time |\n|---|---:|---:|---:|\n")
        for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{k[:100]}` | {n} | {ms:.3f} | {100 * ms / total:.2f}% |\n")
        f.write(f"\nTotal {sum(a[0] for a in agg.values())} launches, {total:.1f} ms.\n")


def full(path, out, command, note):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    with open(out, "w") as f:
        f.write(f"# ncu --set full, {d.get('Kernel Name', ('', '?'))[1]}\n\nCommand: `{command}`\n{note}\n\n| metric | value | unit |\n|---|---:|---|\n")
        for m in FULL_METRICS:
            if m in d:
                f.write(f"| {m} | {d[m][1]} | {d[m][0]} |\n")


if __name__ == "__main__":
    kind, path, out, command = sys.argv[1:5]
    note = sys.argv[5] if len(sys.argv) > 5 else ""
    (launches if kind == "launches" else full)(path, out, command, note)
