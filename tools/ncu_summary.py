#!/usr/bin/env python
"""Summarise ncu artefacts into small text files under profiles/ (run in the dev container).

  python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/rNN_launches.txt
  python tools/ncu_summary.py kernel   gpurun_out/prof.ncu-rep  > profiles/rNN_kernel.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    iname, ival = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[start + 1:]:
        try:
            v = float(r[ival].replace(",", ""))
        except Exception:
            continue
        agg[r[iname][:100]][0] += 1
        agg[r[iname][:100]][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e6:.3f} ms of kernel time "
          "(ncu-serialised, cold cache: compare shares)")
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v / 1e6:12.3f} ms {100 * v / tot:6.2f}%  x{c:<5d} {n}")


def kernel(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"# {path}: {name}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:85s} {vals[i]:>16s} {units[i]}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    if len(rows) > 2:
        hdr = rows[1]
        isrc, isamp = hdr.index("Source"), hdr.index("# Samples")
        data = []
        for r in rows[2:]:
            try:
                data.append((r[isrc], int(r[isamp] or 0)))
            except Exception:
                pass
        tot = sum(d[1] for d in data) or 1
        dm = [i for i, d in enumerate(data) if "DMMA" in d[0]]
        if dm:
            bars = [i for i, d in enumerate(data[:dm[0]]) if "BAR.SYNC" in d[0]]
            ls = (bars[-1] - 10) if bars else dm[0]
            pre = sum(d[1] for d in data[:ls])
            main = sum(d[1] for d in data[ls:dm[-1] + 40])
            post = sum(d[1] for d in data[dm[-1] + 40:])
            print(f"# warp-sample shares: prologue {100 * pre / tot:.1f}%  main loop {100 * main / tot:.1f}%  "
                  f"epilogue {100 * post / tot:.1f}%")
        print("# hottest instructions (samples):")
        for s, n in sorted(data, key=lambda d: -d[1])[:12]:
            print(f"{n:9d} {100 * n / tot:5.1f}%  {s[:90]}")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
