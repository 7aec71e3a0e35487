#!/usr/bin/env python
"""
bench.py -- throughput of the VARDEN advection + MAC-projection hot path on B200.

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--config 3|2|4|5] [--ratio R]

A "step" is one pass of the path advance_timestep.f90:95-124 (advance_premac -> macproject -> scalar_advance ->
make_at_halftime -> velocity_advance) over the synthetic density-stratified Rayleigh-Taylor state of SURVEY 8(d).
Default workload at EVERY N = BASELINE.json configs[2], the configuration the metric / north_star target is quoted on:
3-D 512^3 single level, eight 256^3 reference boxes (max_grid_size = 256, initialize.f90:199), STRONG scaling: 8/4/2/1 boxes per
GPU at 1/2/4/8 GPUs (it fits one B200).  --config 2: 256^3 on one GPU / one 256^3 box per GPU (weak); --config 4: 512^3 per GPU (weak,
1024^3 on 8); --config 5: config 3 at density ratio 1000:1.
Metric = Gcell-updates/s (cells x steps / device time).  Prints ONE JSON line (see the task contract) with
  value     : inputs resident in HBM, device time by CUDA events on the library's stream (max over ranks)
  e2e       : through the C ABI from pinned HOST buffers: H2D of uold/sold/gp/ext forces + step + D2H of unew/snew/rhohalf
  roofline  : dominant kernel family: algorithmic bytes / its CUDA-event time, vs MEASURED_PEAKS.json
  phases    : device ms per step of the reference's own timers (advance_timestep.f90:160-164): MAC / Scalar / Velocity
  nvlink    : bytes this step sent to other ranks / (900 GB/s per direction per GPU) vs the time the exchanges took
  cpu_baseline : the CPU oracle (C restatement, OpenMP, all host cores, -O3 -march=native timing build) on a bounded sample
--impl reference times that CPU restatement alone (the reference Fortran cannot be built here: no Fortran compiler,
FBoxLib absent) and prints the same line with "impl": "reference".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Gcell-updates/s per step (advect+MAC proj)"
UNIT = "Gcell-updates/s"


def ncu_traffic(kernel_family):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of a kernel family from the committed ncu --set full
    capture (profiles/ncu_traffic.json, written by profiles/summarize_ncu.py traffic ...); None if that family was not captured"""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(p)).get(kernel_family, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons during the timed region"""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class CpuOracle:
    """the CPU oracle (C/OpenMP restatement of the reference loops, all host cores, timing build: -O3 -march=native, contraction on)
    on a bounded sample: the same RT problem at n_sample^3 (one 256^3 reference box of the workload by default)"""

    def __init__(self, n_sample, ratio):
        from oracle import oracle as O
        self.O = O
        self.threads = O.use_timing_build(host_cores())      # sets the OpenMP thread count explicitly (torchrun exports OMP_NUM_THREADS=1)
        self.geom, self.P, self.st, self.dt = O.rt_state(n_sample, dim=3, max_grid_size=256, ratio=ratio)
        self.cycles = 0

    def step(self):
        t0 = time.perf_counter()
        out = self.O.advance(self.geom, self.P, self.st, self.dt)
        self.cycles = out["mac_cycles"]
        return time.perf_counter() - t0


def workload(args, world):
    """(cells per direction of a reference box, global grid, scaling, density ratio, label) of the selected BASELINE config"""
    cfg = args.config
    ratio = args.ratio if args.ratio else (1000.0 if cfg == 5 else 2.0)
    n = args.n
    pgrid = {1: [1, 1, 1], 2: [1, 1, 2], 4: [1, 2, 2], 8: [2, 2, 2]}.get(world)
    if pgrid is None:
        raise SystemExit("bench.py: --gpus must be 1, 2, 4 or 8")
    if cfg in (3, 5):
        g = args.global_n or 2 * n
        return n, [g] * 3, "strong", ratio, pgrid, "BASELINE configs[%d]: 3D %d^3 single-level, %d^3 reference boxes, strong scaling" % (cfg - 1, g, n)
    if cfg == 2:
        return n, [n * pgrid[d] for d in range(3)], "weak", ratio, pgrid, "BASELINE configs[1]: 3D %d^3 per GPU (one reference box), weak scaling" % n
    if cfg == 4:
        return n, [2 * n * pgrid[d] for d in range(3)], "weak", ratio, pgrid, "BASELINE configs[3]: 3D %d^3 per GPU (eight %d^3 boxes), weak scaling" % (2 * n, n)
    raise SystemExit("bench.py: --config must be 2, 3, 4 or 5")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    n, nglob, scaling, ratio, pgrid, label = workload(args, max(world, 1) if world in (1, 2, 4, 8) else 1)
    ns = args.cpu_n
    cpu = CpuOracle(ns, ratio)
    for _ in range(max(min(args.warmup, 1), 1)):
        cpu.step()                                  # untimed: page-in, thread pool
    times = [cpu.step() for _ in range(args.steps)]
    t_tot = sum(times)
    value = cpu.geom.ncells * args.steps / t_tot / 1e9
    same = [ns] * 3 == list(nglob)
    sample = "same RT problem at %d^3 = %s of the %dx%dx%d workload (one pass per step, %.1f s each), %d V-cycles" % (
        ns, "all" if same else "1/%d" % round(nglob[0] * nglob[1] * nglob[2] / ns ** 3), nglob[0], nglob[1], nglob[2], t_tot / max(args.steps, 1), cpu.cycles)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_tot / max(args.steps, 1), "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "same_config": same,
            "config": {"workload": label + " (RT ratio %g:1); CPU arm on a bounded sample: %d^3" % (ratio, ns)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.threads, "kind": "port",
                             "sample": sample + "; C/OpenMP restatement of the reference loops (oracle/, -O3 -march=native build), not the Fortran binary"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", type=int, default=3, help="BASELINE.json config: 3 (default: 512^3 strong scaling, eight 256^3 boxes), 2 (256^3 per GPU, weak), "
                                                          "4 (512^3 per GPU, weak), 5 (config 3 at ratio 1000:1)")
    ap.add_argument("--n", type=int, default=256, help="cells per direction of a reference box (max_grid_size)")
    ap.add_argument("--ratio", type=float, default=0.0, help="density ratio (default 2, config 5: 1000)")
    ap.add_argument("--cpu-n", type=int, default=256, help="grid size of the bounded CPU sample (256^3 = one reference box: a few seconds per pass)")
    ap.add_argument("--force-nccl", action="store_true", help="A/B: ghost exchanges through NCCL send/recv instead of the peer-memory transport")
    ap.add_argument("--xchg", default="push", choices=["push", "nccl", "pull", "pushk"],
                    help="A/B: ghost exchange of the fused multigrid levels -- push: stored by the sweep kernel itself (default); nccl; pull / pushk: a pull kernel "
                         "before / a push kernel after every sweep")
    ap.add_argument("--fuse-min", type=int, default=0, help="A/B: smallest multigrid level the fused smoother runs on (library default: 128, 64 on rank-split levels)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--global-n", type=int, default=0, help="configs 3/5: global cells per direction (default 2 x --n)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import ctypes as C
    import varden_b200 as V
    from varden_b200.problems import rt_problem, mf_alloc

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        print("bench.py: WORLD_SIZE (%d) != --gpus (%d); launch with torchrun for N>1" % (world, args.gpus), file=sys.stderr)
        if args.gpus > 1 and world == 1:
            sys.exit(2)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n, nglob, scaling, ratio, pgrid, label = workload(args, world)
    if any(nglob[d] % (n * pgrid[d]) for d in range(3)):
        raise SystemExit("bench.py: the process grid %s does not divide the %s grid of %d^3 boxes" % (pgrid, nglob, n))
    args.ratio = ratio
    phi = [float(nglob[d]) / n for d in range(3)]            # domain [0, nglob/n]^3: dx = 1/n in every configuration
    mgs = n
    from varden_b200.problems import Geom, PERIODIC, NO_SLIP_WALL
    from varden_b200 import parallel as PAR
    gfull = Geom(3, nglob, [[PERIODIC, PERIODIC], [PERIODIC, PERIODIC], [NO_SLIP_WALL, NO_SLIP_WALL]],
                 prob_hi=phi, max_grid_size=mgs)
    ids, rlo, rhi, pg = PAR.partition(gfull, world)
    geom, st, dt = rt_problem(nglob, dim=3, max_grid_size=mgs, ratio=args.ratio, prob_hi=phi, box_ids=ids[rank])
    ncells_global = gfull.ncells
    dim, nscal = 3, 2
    prm = V.default_params()
    ctx = V.Context(3, geom.boxes, gfull.dlo, gfull.dhi, gfull.phys_bc, gfull.dx, params=prm, device=local)
    if world > 1:
        if args.force_nccl:
            args.xchg = "nccl"
        ctx.comm_tune({"push": 0, "nccl": 1, "pull": 2, "pushk": 3}[args.xchg])
        PAR.init_comm(ctx, rank, world, rlo, rhi)

    if args.fuse_min:
        ctx.mg_tune(args.fuse_min, -1)

    # ---- pinned host buffers = what the Fortran driver would hand over (ghosts filled as varden.f90:291-300) ----
    spec_in = [("UOLD", "uold", 3, dim), ("SOLD", "sold", 3, nscal), ("GP", "gp", 1, dim),
               ("EXT_VEL_FORCE", "ext_vel_force", 1, dim), ("EXT_SCAL_FORCE", "ext_scal_force", 1, nscal)]
    spec_out = [("UNEW", 3, dim), ("SNEW", 3, nscal), ("RHOHALF", 1, 1)]

    def pinned_like(a):
        t = torch.empty(a.size, dtype=torch.float64).pin_memory()
        v = t.numpy().reshape(a.shape, order='F')
        return t, v

    host_in, host_out, keep = {}, {}, []
    for fld, key, ng, nc in spec_in:
        host_in[fld] = []
        for a in st[key]:
            t, v = pinned_like(a); v[...] = a; keep.append(t); host_in[fld].append(v)
        ctx.upload_mf(fld, host_in[fld], ng, nc)
    ctx.fill_and_physbc("UOLD", 0)
    ctx.fill_and_physbc("SOLD", dim)
    ctx.fill_boundary("GP")
    for fld, key, ng, nc in spec_in[:3]:
        ctx.download_mf(fld, host_in[fld], ng, nc)          # host copies now carry the path-boundary ghost cells
    for fld, ng, nc in spec_out:
        host_out[fld] = []
        for ib in range(geom.nboxes):
            t, v = pinned_like(np.empty(geom.box_shape(ib, ng) + (nc,), order='F')); keep.append(t); host_out[fld].append(v)
    h2d = sum(v.nbytes for fld in host_in for v in host_in[fld])
    d2h = sum(v.nbytes for fld in host_out for v in host_out[fld])

    def step_resident():
        return ctx.advance(dt)

    hs = ctx.host_state(uold=host_in["UOLD"], sold=host_in["SOLD"], gp=host_in["GP"], ext_vel_force=host_in["EXT_VEL_FORCE"],
                        ext_scal_force=host_in["EXT_SCAL_FORCE"], unew=host_out["UNEW"], snew=host_out["SNEW"], rhohalf=host_out["RHOHALF"])

    def step_e2e():
        # the reference-facing call with HOST multifabs: H2D of the five inputs, the step, D2H of the three outputs
        return ctx.advance_host(dt, hs)

    # ---- resident timing ----
    for _ in range(max(args.warmup, 3)):
        cyc, res = step_resident()
    ctx.sync()
    l0 = ctx.launch_count()
    cb0 = ctx.comm_bytes()
    sampler = ClockSampler(local); sampler.start()
    xs = torch.cuda.ExternalStream(ctx.stream_ptr())        # the library's launching stream
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    barrier()
    t0 = time.perf_counter()
    ev0.record(xs)
    for _ in range(args.steps):
        cyc, res = step_resident()
    ev1.record(xs)
    barrier()
    wall_host = time.perf_counter() - t0
    wall = max_over_ranks(ev0.elapsed_time(ev1) / 1e3)       # device time by CUDA events on the launching stream, max over ranks
    launches = ctx.launch_count() - l0
    comm_bytes = (ctx.comm_bytes() - cb0) / args.steps
    # ---- the same K steps once more with the library's per-launch CUDA events on: the per-kernel-family table, the phases and the roofline
    # (kept out of the pass above so that `value` does not pay two event records per launch) ----
    ctx.prof_enable(True)
    if world > 1:
        ctx.debug_counters()                                 # switch the flag-wait accounting of the fused smoother on
    barrier()
    evp0, evp1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evp0.record(xs)
    for _ in range(args.steps):
        cyc, res = step_resident()
    evp1.record(xs)
    barrier()
    wall_prof = max_over_ranks(evp0.elapsed_time(evp1) / 1e3)
    prof = ctx.prof_report()
    ctx.prof_enable(False)
    waits, by_rank = None, None
    if world > 1:
        # per rank: time per kernel family (the table below is rank 0's) and what the boundary CTAs of the fused sweeps spent waiting for a neighbour
        mine = {"ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items() if v["ms"] > 0 and not k.startswith("phase:")},
                "flag_wait": {k: {"ms_summed_over_ctas_per_step": v[0] / args.steps, "longest_us": v[1], "waits_per_step": v[2] / args.steps}
                              for k, v in ctx.debug_counters().items() if v[2]}}
        by_rank = [None] * world
        dist.all_gather_object(by_rank, mine)
    phases = {k[6:]: v["ms"] / args.steps for k, v in prof.items() if k.startswith("phase:")}
    prof = {k: v for k, v in prof.items() if not k.startswith("phase:")}
    dev_ms = sum(p["ms"] for p in prof.values())
    # one in-order stream; the host only syncs for the per-V-cycle residual norm, so event time ~ host wall time
    ms_per_step = 1e3 * wall / args.steps
    value = ncells_global * args.steps / wall / 1e9

    # ---- e2e timing (host buffers, copies inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        for _ in range(2):
            step_e2e()
        barrier()
        # vdn_advance_host returns when the outputs are complete on the host, so host wall time IS the end-to-end time
        # (copies run on the library's own copy streams; an event pair on the compute stream would miss the last D2H)
        t_h = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        barrier()
        t_e = max_over_ranks(time.perf_counter() - t_h)
        e2e = {"value": ncells_global * args.steps / t_e / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(d2h) * world,
               "ms_per_step": 1e3 * t_e / args.steps,
               "how": "vdn_advance_host (C ABI, pinned host multifabs in and out, copies on the library's copy streams overlapped with the stages); "
                      "host clock around the blocking calls, max over ranks"}
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- roofline of the dominant kernel family ----
    peak, peak_src = measured_peak()
    fam = {}
    for name, p in prof.items():
        if p["ms"] <= 0:
            continue
        fam[name] = {"launches": p["launches"], "ms_total": p["ms"], "share": p["ms"] / dev_ms if dev_ms else None,
                     "alg_gb": p["alg_bytes"] / 1e9, "gbs": (p["alg_bytes"] / 1e9) / (p["ms"] / 1e3) if p["alg_bytes"] > 0 else None}
        if fam[name]["gbs"] is not None:
            fam[name]["frac"] = fam[name]["gbs"] / peak
    # the dominant KERNEL: the four level-0 families mg_wave_{smooth,pro,down,up}_l0 are instantiations of one kernel (k_sweep3: a GSRB
    # sweep [+ prolongation] [+ residual + restriction / norm]) and are judged together; every other family is one kernel
    groups = {}
    for name, t in fam.items():
        g = "k_sweep3 (mg_wave_*_l0: fused GSRB sweep +prolong/+residual+restrict/+norm, level 0)" if name.startswith("mg_wave_") and name.endswith("_l0") else name
        groups.setdefault(g, []).append(name)
    gtime = {g: sum(fam[n]["ms_total"] for n in ns) for g, ns in groups.items()}
    top = max(gtime, key=gtime.get) if gtime else None
    roof = None
    if top:
        ns = groups[top]
        ms = gtime[top]
        alg = sum(fam[n]["alg_gb"] for n in ns) * 1e9
        nl = sum(fam[n]["launches"] for n in ns)
        gbs = alg / 1e9 / (ms / 1e3) if alg > 0 else None
        tr = [(ncu_traffic(n), fam[n]["launches"]) for n in ns]
        traffic = sum(b * l for b, l in tr) / nl if all(b is not None for b, _ in tr) and nl else None
        roof = {"bound": "hbm", "kernel": top, "families": ns, "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": (gbs / peak) if gbs else None,
                "traffic": traffic, "alg_bytes_per_launch": alg / max(nl, 1), "peak_source": peak_src, "avg_launch_ms": ms / max(nl, 1),
                "share_of_step": ms / dev_ms if dev_ms else None,
                "note": "achieved = SURVEY 8(a) algorithmic bytes (80 B/cell per sweep = two 40 B/cell colour half-sweeps, +17 prolongation, +48 residual, "
                        "+9 restriction) / CUDA-event time; traffic = DRAM bytes per launch from profiles/ncu_traffic.json (ncu --set full), launch-weighted"}

    cpu = None
    if not args.no_cpu and world == 1:
        co = CpuOracle(args.cpu_n, args.ratio)
        co.step()                                   # untimed: page-in, thread pool
        t = co.step()
        cpu = {"value": co.geom.ncells / t / 1e9, "unit": UNIT, "cores": co.threads, "kind": "port",
               "sample": "same RT problem at %d^3 (1/%d of the workload), one pass (%.1f s, %d V-cycles); C/OpenMP restatement of the reference loops "
                         "(-O3 -march=native build), not the Fortran binary" % (args.cpu_n, round(ncells_global / args.cpu_n ** 3), t, co.cycles)}
    # NVLink: what this rank sent to other ranks per step against 900 GB/s per direction per GPU, and against the time the exchanges took
    nvlink = None
    if world > 1:
        t_x = sum(fam[k]["ms_total"] for k in fam if k in ("mg_halo_exchange", "halo_exchange", "mg_agglomerate")) / args.steps
        nvlink = {"bytes_sent_per_step_per_gpu": comm_bytes, "peak_gbs_per_direction": 900.0, "wire_ms_at_peak": comm_bytes / 900e9 * 1e3,
                  "exchange_ms_per_step": t_x, "achieved_gbs": (comm_bytes / 1e9) / (t_x / 1e3) if t_x > 0 else None,
                  "frac": (comm_bytes / 900e9 * 1e3) / t_x if t_x > 0 else None,
                  "note": "latency-bound: the exchanges move surface data only; frac = wire time at peak / measured exchange time"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": label + "; variable-density RT (ratio %g:1), periodic x,y / no-slip z, nscal=2, slope_order=4, MAC rel tol 1e-10, "
                                   "%d reference box(es) of %d^3 per GPU, global %dx%dx%d; the same input state every step" % (args.ratio, geom.nboxes, n, nglob[0], nglob[1], nglob[2]),
                       "parallelism": ("1 region per GPU, process grid %s, %s ghost exchanges, NCCL allreduce, coarse MG levels agglomerated"
                                       % (pgrid, {"nccl": "NCCL send/recv", "push": "peer-memory (CUDA-IPC; fused MG sweeps push their boundary results)", "pull": "peer-memory (pull kernel per exchange)", "pushk": "peer-memory (push kernel per MG sweep)"}[args.xchg])) if world > 1 else "single GPU",
                       "l2_policy": "inputs (%.1f GB of fields per GPU) exceed the 126 MB L2; no explicit flush" % (45 * 8 * geom.nboxes * (n + 6) ** 3 / 1e9),
                       "mac_vcycles_per_step": cyc, "mac_resnorm": res, "kernel_ms_sum_per_step": dev_ms / args.steps,
                       "ms_per_step_with_per_launch_events": 1e3 * wall_prof / args.steps,
                       "timing": "value / ms_per_step: K steps, CUDA events on the library's stream, no per-launch events; kernels / phases / roofline: "
                                 "the same K steps run once more with the library's per-launch CUDA events on",
                       "host_wall_ms_per_step": 1e3 * wall_host / args.steps},
            "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
            "phases_ms_per_step": phases, "nvlink": nvlink,
            "kernels": fam, "by_rank": by_rank}
    if rank == 0:
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
