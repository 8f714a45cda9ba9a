#!/usr/bin/env python
"""Summarise ncu output brought back in gpurun_out/ into small text files for profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv STEPS > profiles/rN_launches.txt
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep > profiles/rN_full.txt
"""
import collections
import csv
import subprocess
import sys


def launches(path, steps):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        a = agg.setdefault(r[ki][:90], [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none, {steps} step(s); cold-cache serialised: compare SHARES")
    print(f"# total {tot / steps:.1f} us/step over {sum(a[0] for a in agg.values()) // steps} launches/step")
    for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{v / steps:10.1f} us/step {n / steps:5.1f} launches/step {100 * v / tot:5.1f}%  {k}")


WANT = ["Kernel Name", "launch__grid_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "lts__t_sector_hit_rate.pct"]


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [(w, hdr.index(w)) for w in WANT if w in hdr]
    print("# ncu --set full --clock-control none; one row per captured launch")
    print(" | ".join(f"{w} [{units[i]}]" for w, i in cols))
    for r in rows[2:]:
        print(" | ".join(r[i][:60] for _, i in cols))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]))
    else:
        full(sys.argv[2])
