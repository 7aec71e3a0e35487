#!/usr/bin/env python3
"""
profiles/summarize_ncu.py -- turn ncu output brought back in gpurun_out/ into the small text summaries kept under profiles/.

  launch list : ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv <cmd>
                -> per-kernel launches / total time / share of the captured launches  (cold-cache, serialised: shares only)
  full capture: ncu --set full ... -o X ; ncu -i X.ncu-rep --page raw --csv > X_raw.csv
                -> per-launch duration, DRAM bytes read+written, DRAM throughput %, registers, occupancy, L2 hit rate
"""
import collections
import csv
import sys


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    h = rows[0]
    ki, vi, ui, mi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("Metric Name"), h.index("Grid Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        t = float(r[vi].replace(",", ""))
        t_us = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}[r[ui]] * t
        a = agg.setdefault(r[ki], [0, 0.0, 0.0])
        a[0] += 1
        a[1] += t_us
        a[2] = max(a[2], t_us)
    tot = sum(a[1] for a in agg.values())
    n = sum(a[0] for a in agg.values())
    out.write("# %s: %d launches, %.1f ms of kernel time (ncu-serialised, cold cache: compare SHARES only)\n" % (path, n, tot / 1e3))
    out.write("# launches   total_us   share   max_us  kernel\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.write("%8d %11.1f %7.4f %8.1f  %s\n" % (a[0], a[1], a[1] / tot, a[2], k))


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]


def full(path, out):
    rows = list(csv.reader(open(path)))
    h, u = rows[0], rows[1]
    out.write("# %s (ncu --set full, one block per captured launch)\n" % path)
    for r in rows[2:]:
        out.write("kernel: %s\n" % r[h.index("Kernel Name")])
        for w in WANT:
            if w in h:
                out.write("    %-70s %s %s\n" % (w, r[h.index(w)], u[h.index(w)]))


if __name__ == "__main__":
    mode, src = sys.argv[1], sys.argv[2]
    dst = open(sys.argv[3], "w") if len(sys.argv) > 3 else sys.stdout
    (launches if mode == "launches" else full)(src, dst)
