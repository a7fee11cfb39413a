#!/usr/bin/env python
"""Turn ncu artefacts brought back in gpurun_out/ into the small text summaries kept under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/r01_bench_launches.csv > profiles/r01_bench_launches.txt
    python tools/summarize_ncu.py full gpurun_out/r01_env128.ncu-rep           > profiles/r01_env128.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_active.avg", "smsp__cycles_active.avg",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if r and r[0] == "ID":
            hdr, data = r, rows[i + 1:]
            break
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) > vi:
            agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
    unit = data[0][ui]
    total = sum(sum(v) for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none ({path}); cold-cache, serialised launches")
    print(f"# {'count':>5} {'avg_' + unit:>12} {'min':>10} {'max':>10} {'share':>7}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"  {len(v):5d} {sum(v) / len(v):12.1f} {min(v):10.1f} {max(v):10.1f} {sum(v) / total:7.1%}  {k[:110]}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    print(f"# ncu --set full --clock-control none --import-source on ({path}); {len(data)} launches")
    for r in data:
        print(f"# kernel: {r[kn][:120]}")
        break
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:82s} {units[i]:14s} " + "  ".join(r[i] for r in data))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
