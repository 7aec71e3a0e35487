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


def family(kname, grid):
    """bench.py kernel-family name of a captured launch (level-0 launches of the 256^3 bench are the ones with the large grid)"""
    import re
    m = re.search(r"k_(sweep3|sweep2|sweep|wave)<([^>]*)>", kname)
    if m:
        a = [x.strip() for x in m.group(2).split(",")]
        pre, post = (int(a[0]), int(a[1])) if m.group(1) in ("sweep2", "sweep3") else (int(a[1]), int(a[2]))
        return {2: "mg_wave_down_l0", 3: "mg_wave_up_l0"}.get(post, "mg_wave_pro_l0" if pre else "mg_wave_smooth_l0")
    if "k_velpred_march" in kname:
        return "velpred"
    m = re.search(r"k_mkflux_march<\s*\d+\s*,\s*(\d+)", kname)
    if m:       # <NC, CONSMASK, ...>: the conservative (density) launch is a scalar one; non-conservative launches are velocity components or the tracer
        return "mkflux_scal" if int(m.group(1)) else "mkflux_vel"
    for k, f in (("k_gsrb", "mg_gsrb_l0"), ("k_residual", "mg_residual_l0"), ("k_mf_normal3", "mf_normal3"), ("k_mf_trans6", "mf_trans6"),
                 ("k_mf_final3", "mf_final3"), ("k_vp_normal3", "vp_normal3"), ("k_vp_trans6", "vp_trans6"), ("k_vp_final3", "vp_final3"),
                 ("k_mkvelforce", "mkvelforce")):
        if k in kname:
            return f
    if "k_update" in kname:
        return "update_vel" if ", 3, 1>" in kname.replace("(bool)1", "1") or "3, 3, true" in kname else "update_scal"
    return None


def traffic(path, out):
    """{family: {dram_bytes_per_launch, duration_us, kernel}} from a --set full raw CSV: the LARGEST launch of every family"""
    import json
    rows = list(csv.reader(open(path)))
    h, u = rows[0], rows[1]
    ik, ir, iw, it, ig = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum"), h.index("Grid Size")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tsc = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
    res = {}
    for r in rows[2:]:
        f = family(r[ik], r[ig])
        if not f:
            continue
        b = float(r[ir].replace(",", "")) * scale[u[ir]] + float(r[iw].replace(",", "")) * scale[u[iw]]
        t = float(r[it].replace(",", "")) * tsc[u[it]]
        if f not in res or b > res[f]["dram_bytes_per_launch"]:
            res[f] = {"dram_bytes_per_launch": b, "duration_us": t, "kernel": r[ik], "grid": r[ig], "source": path}
    json.dump(res, out, indent=1, sort_keys=True)
    out.write("\n")


if __name__ == "__main__":
    mode, src = sys.argv[1], sys.argv[2]
    dst = open(sys.argv[3], "w") if len(sys.argv) > 3 else sys.stdout
    {"launches": launches, "full": full, "traffic": traffic}[mode](src, dst)
