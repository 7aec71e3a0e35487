"""
GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI, against the CPU oracle
on the same seeded inputs.  Tolerances are the ones BASELINE.json:north_star states:
  * edge states / updated scalars and velocities: 1e-12 relative L-inf (stage-wise, identical inputs) -- in practice
    the kernels are bit-exact with the oracle (same operation order, FMA contraction off on both sides);
  * projected MAC velocity and phi: 10x the MAC solver tolerance = 1e-9 relative;
  * all fields after 10 steps: 1e-8 relative.
"""
import numpy as np
import pytest

from oracle import oracle as O
from util import make_ctx, upload_state, relerr, download_like

pytestmark = pytest.mark.gpu

TOL_EDGE = 1e-12
TOL_MAC = 1e-9
TOL_10STEP = 1e-8

W, NS, IN, OUT, PER, SYM = O.SLIP_WALL, O.NO_SLIP_WALL, O.INLET, O.OUTLET, O.PERIODIC, O.SYMMETRY

CASES = {
    "rt3d_1box":   lambda: O.rt_state(32, dim=3, max_grid_size=32),
    "rt3d_8box":   lambda: O.rt_state(32, dim=3, max_grid_size=16),
    "rt3d_aniso":  lambda: O.rt_state([32, 16, 24], dim=3, max_grid_size=16),
    "rand3d_slip": lambda: O.random_state(16, dim=3, max_grid_size=16, phys_bc=[[W, W], [W, W], [W, W]], seed=1),
    "rand3d_mixed": lambda: O.random_state([16, 12, 20], dim=3, max_grid_size=32, phys_bc=[[IN, OUT], [NS, W], [PER, PER]], seed=2),
    "rand3d_outx": lambda: O.random_state(16, dim=3, max_grid_size=8, phys_bc=[[OUT, IN], [PER, PER], [W, OUT]], seed=3),
    "rand3d_per":  lambda: O.random_state(16, dim=3, max_grid_size=8, phys_bc=[[PER, PER]] * 3, seed=4),
    "rt2d_4box":   lambda: O.rt_state(64, dim=2, max_grid_size=32),
    "rand2d_mixed": lambda: O.random_state([24, 16], dim=2, max_grid_size=32, phys_bc=[[IN, OUT], [NS, W]], seed=5),
    "rand2d_walls": lambda: O.random_state(32, dim=2, max_grid_size=16, phys_bc=[[NS, NS], [NS, NS]], seed=6),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_stagewise_parity(case):
    geom, P, st, dt = CASES[case]()
    dim, nscal = geom.dim, P.nscal
    ref = O.stagewise(geom, P, st, dt, mac_rel_eps=1e-13)
    ctx = make_ctx(geom, P)
    upload_state(ctx, geom, P, st)
    errs = {}

    # advance_premac: mkvelforce + velpred
    ctx.mkvelforce("SOLD", 1.0)
    errs["vel_force"] = relerr(geom, download_like(ctx, geom, "VEL_FORCE", ref["vel_force_1"], 1, dim), ref["vel_force_1"], 1)
    ctx.velpred(dt)
    for d in range(dim):
        got = download_like(ctx, geom, "UMAC_" + "XYZ"[d], ref["umac_pred"][d], 1, 1)
        errs["umac_pred%d" % d] = relerr(geom, got, ref["umac_pred"][d], 1, d)

    # MAC projection from the same predicted velocities
    ncyc, res = ctx.macproject(rel_eps=1e-10)
    assert res <= 1e-10
    scale = max(max(np.abs(O.valid(geom, a, ib, 1, d)).max() for ib, a in enumerate(ref["umac"][d])) for d in range(dim))
    for d in range(dim):
        got = download_like(ctx, geom, "UMAC_" + "XYZ"[d], ref["umac"][d], 1, 1)
        e = max(np.abs(O.valid(geom, g, ib, 1, d) - O.valid(geom, r, ib, 1, d)).max() for ib, (g, r) in enumerate(zip(got, ref["umac"][d])))
        errs["umac_proj%d" % d] = e / scale
        assert e / scale <= TOL_MAC, (case, d, e / scale)
    assert ctx.divumac() <= 1e-9 * max(1.0, scale / min(geom.dx[:dim]))

    # downstream stages on the ORACLE's projected umac (identical inputs => 1e-12 bar)
    for d in range(dim):
        ctx.upload_mf("UMAC_" + "XYZ"[d], ref["umac"][d], 1, 1)
    ctx.mkscalforce(1.0)
    ctx.mkflux(False, dt)
    for d in range(dim):
        got = download_like(ctx, geom, "SEDGE_" + "XYZ"[d], ref["sedge"][d], 0, nscal)
        errs["sedge%d" % d] = relerr(geom, got, ref["sedge"][d], 0, d)
        gotf = download_like(ctx, geom, "SFLUX_" + "XYZ"[d], [a[..., :1].copy(order='F') for a in ref["sflux"][d]], 0, 1)
        errs["sflux%d" % d] = relerr(geom, gotf, [a[..., :1] for a in ref["sflux"][d]], 0, d)
    ctx.mkscalforce(0.0)
    ctx.update(False, dt)
    errs["snew"] = relerr(geom, download_like(ctx, geom, "SNEW", ref["snew"], 3, nscal), ref["snew"], 3)
    ctx.make_at_halftime()
    errs["rhohalf"] = relerr(geom, download_like(ctx, geom, "RHOHALF", ref["rhohalf"], 1, 1), ref["rhohalf"], 1)
    ctx.mkvelforce("SOLD", 1.0)
    ctx.mkflux(True, dt)
    for d in range(dim):
        got = download_like(ctx, geom, "UEDGE_" + "XYZ"[d], ref["uedge"][d], 0, dim)
        errs["uedge%d" % d] = relerr(geom, got, ref["uedge"][d], 0, d)
    ctx.mkvelforce("RHOHALF", 0.0)
    errs["vel_force_2"] = relerr(geom, download_like(ctx, geom, "VEL_FORCE", ref["vel_force_2"], 1, dim), ref["vel_force_2"], 1)
    ctx.update(True, dt)
    errs["unew"] = relerr(geom, download_like(ctx, geom, "UNEW", ref["unew"], 3, dim), ref["unew"], 3)
    # ghost cells of every box, inter-box ghosts included (what ml_restrict_and_fill leaves behind, update.f90:103-107; hgproject's
    # create_uvec reads rhohalf(lo-1:hi+1)): a download carries the box's whole ghosted extent
    errs["snew_ghost"] = relerr(geom, download_like(ctx, geom, "SNEW", ref["snew"], 3, nscal), ref["snew"], 3, full=True)
    errs["unew_ghost"] = relerr(geom, download_like(ctx, geom, "UNEW", ref["unew"], 3, dim), ref["unew"], 3, full=True)
    errs["rhohalf_ghost"] = relerr(geom, download_like(ctx, geom, "RHOHALF", ref["rhohalf"], 1, 1), ref["rhohalf"], 1, full=True)
    ctx.close()
    print(case, {k: "%.2e" % v for k, v in errs.items()}, "vcycles", ncyc)
    bad = {k: v for k, v in errs.items() if not k.startswith("umac_proj") and not (v <= TOL_EDGE)}
    assert not bad, (case, bad)


@pytest.mark.parametrize("case", ["rt3d_8box", "rand3d_mixed", "rt2d_4box"])
def test_one_step_end_to_end(case):
    """whole path through vdn_advance, both MAC solves converged to 1e-13 (SURVEY Q10 mode b)"""
    geom, P, st, dt = CASES[case]()
    dim, nscal = geom.dim, P.nscal
    ref = O.advance(geom, P, st, dt, mac_rel_eps=1e-13)
    ctx = make_ctx(geom, P)
    upload_state(ctx, geom, P, st)
    ncyc, res = ctx.advance(dt, mac_rel_eps=1e-13)
    e_s = relerr(geom, download_like(ctx, geom, "SNEW", ref["snew"], 3, nscal), ref["snew"], 3)
    e_u = relerr(geom, download_like(ctx, geom, "UNEW", ref["unew"], 3, dim), ref["unew"], 3)
    e_r = relerr(geom, download_like(ctx, geom, "RHOHALF", ref["rhohalf"], 1, 1), ref["rhohalf"], 1)
    ctx.close()
    print(case, "snew %.2e unew %.2e rhohalf %.2e" % (e_s, e_u, e_r), "vcycles", ncyc, res)
    assert e_s <= 1e-10 and e_u <= 1e-10 and e_r <= 1e-10


@pytest.mark.parametrize("case", ["rt3d_1box", "rt3d_8box", "rt2d_4box"])
def test_advance_host_pipelined(case):
    """vdn_advance_host (host multifabs in and out, copies overlapped with the stages on separate streams) gives the same
    fields as upload + vdn_advance + download -- twice in a row, so a stale event or an early D2H would show"""
    geom, P, st, dt = CASES[case]()
    dim, nscal = geom.dim, P.nscal
    ref = O.advance(geom, P, st, dt, mac_rel_eps=1e-13)
    ctx = make_ctx(geom, P)
    for rep in range(2):
        out = dict(unew=[np.full_like(a, np.nan) for a in ref["unew"]], snew=[np.full_like(a, np.nan) for a in ref["snew"]],
                   rhohalf=[np.full_like(a, np.nan) for a in ref["rhohalf"]])
        hs = ctx.host_state(uold=st["uold"], sold=st["sold"], gp=st["gp"], ext_vel_force=st["ext_vel_force"],
                            ext_scal_force=st["ext_scal_force"], **out)
        ctx.advance_host(dt, hs, mac_rel_eps=1e-13)
        e_s = relerr(geom, out["snew"], ref["snew"], 3, full=True)          # ghost cells of every box included
        e_u = relerr(geom, out["unew"], ref["unew"], 3, full=True)
        e_r = relerr(geom, out["rhohalf"], ref["rhohalf"], 1, full=True)
        print(case, rep, "snew %.2e unew %.2e rhohalf %.2e" % (e_s, e_u, e_r))
        assert e_s <= 1e-10 and e_u <= 1e-10 and e_r <= 1e-10
    ctx.close()


def test_ten_steps():
    """all fields within 1e-8 relative after 10 steps at the reference's MAC tolerance (1e-10)"""
    geom, P, st, dt = O.rt_state(32, dim=3, max_grid_size=16)
    dt = 0.2 * dt          # hgproject is outside the path: keep the un-projected, gravity-accelerated velocity within CFL for 10 steps
    dim, nscal = geom.dim, P.nscal
    ctx = make_ctx(geom, P)
    upload_state(ctx, geom, P, st)
    s_ref = dict(st)
    for step in range(10):
        ref = O.advance(geom, P, s_ref, dt)
        ctx.advance(dt)
        # next step: (uold, sold) <- (unew, snew); hgproject stays outside the path in both arms
        s_ref = dict(s_ref, uold=ref["unew"], sold=ref["snew"])
        un = download_like(ctx, geom, "UNEW", ref["unew"], 3, dim)
        sn = download_like(ctx, geom, "SNEW", ref["snew"], 3, nscal)
        ctx.upload_mf("UOLD", un, 3, dim)
        ctx.upload_mf("SOLD", sn, 3, nscal)
    e_u = relerr(geom, un, ref["unew"], 3)
    e_s = relerr(geom, sn, ref["snew"], 3)
    ctx.close()
    ndiff = sum(int((O.valid(geom, a, ib, 3) != O.valid(geom, b, ib, 3)).sum()) for ib, (a, b) in enumerate(zip(un, ref["unew"])))
    print("10 steps: unew %.2e snew %.2e (unew entries differing bitwise: %d)" % (e_u, e_s, ndiff))
    assert e_u <= TOL_10STEP and e_s <= TOL_10STEP


def mass(geom, mf, comp=0):
    """sum of one component over the valid cells of every box (scalars carry 3 ghost cells)"""
    return float(sum(O.valid(geom, a, ib, 3)[..., comp].sum(dtype=np.float64) for ib, a in enumerate(mf)))


def test_mass_helper_on_the_oracle():
    """the helper the full-size test uses, on a case the CPU oracle finishes in a second (runs in the GPU tier next to its user)"""
    geom, P, st, dt = O.rt_state(24, dim=3, max_grid_size=12)
    out = O.advance(geom, P, st, dt)
    assert abs(mass(geom, out["snew"]) - mass(geom, st["sold"])) <= 1e-12 * abs(mass(geom, st["sold"]))


def test_full_size_properties():
    """BASELINE.json config 2 at its full size (3-D 256^3, one box): size-independent properties instead of a 256^3 oracle pass.
      * the MAC projection leaves max|mac_rhs - div(U^MAC)| <= 10 x 1e-10 x its value before (macproject.f90:203-221 prints this norm;
        it equals the multigrid residual, north_star bar: 10x the reference's solver tolerance);
      * the conservative density update telescopes (update.f90:250-253) and the no-slip z faces carry no flux: sum(rho) is constant to
        round-off over the whole step;
      * a uniform state at rest with no forcing is a fixed point: snew == sold bit for bit, unew == 0, zero V-cycles."""
    n = 256
    geom, P, st, dt = O.rt_state(n, dim=3, max_grid_size=n)
    ctx = make_ctx(geom, P)
    upload_state(ctx, geom, P, st)
    ctx.mkvelforce("SOLD", 1.0)
    ctx.velpred(dt)
    pre = ctx.divumac()
    ncyc, res = ctx.macproject(rel_eps=1e-10)
    post = ctx.divumac()
    print("256^3: |rhs - div umac| before %.3e after %.3e (ratio %.2e), %d V-cycles, |r|/|rh| %.2e" % (pre, post, post / pre, ncyc, res))
    assert pre > 0 and res <= 1e-10 and post <= 1e-9 * pre
    ctx.advance(dt)
    snew = [np.full_like(a, np.nan) for a in st["sold"]]
    ctx.download_mf("SNEW", snew, 3, P.nscal)
    m0, m1 = mass(geom, st["sold"]), mass(geom, snew)
    print("256^3: sum(rho) %.15e -> %.15e (relative change %.2e)" % (m0, m1, abs(m1 - m0) / abs(m0)))
    assert np.all(np.isfinite(O.valid(geom, snew[0], 0, 3))) and abs(m1 - m0) <= 1e-12 * abs(m0)
    # fixed point: rho = tracer = 1.5 (exactly representable, so the wall extrapolation reproduces it), everything else zero
    for key, val in (("uold", 0.0), ("sold", 1.5), ("gp", 0.0), ("ext_vel_force", 0.0), ("ext_scal_force", 0.0)):
        for a in st[key]:
            a[...] = val
    upload_state(ctx, geom, P, st)
    ncyc, res = ctx.advance(dt)
    ctx.download_mf("SNEW", snew, 3, P.nscal)
    unew = [np.full_like(a, np.nan) for a in st["uold"]]
    ctx.download_mf("UNEW", unew, 3, geom.dim)
    ctx.close()
    assert ncyc == 0
    assert np.all(O.valid(geom, snew[0], 0, 3) == 1.5)
    assert np.all(O.valid(geom, unew[0], 0, 3) == 0.0)


def test_mg_high_density_ratio():
    """1000:1 density ratio (config 5): the solver must still reach 1e-10 and agree with the oracle to 1e-9"""
    geom, P, st, dt = O.rt_state(32, dim=3, max_grid_size=32, ratio=1000.0)
    ref = O.stagewise(geom, P, st, dt, mac_rel_eps=1e-13)
    ctx = make_ctx(geom, P)
    upload_state(ctx, geom, P, st)
    ctx.mkvelforce("SOLD", 1.0)
    ctx.velpred(dt)
    ncyc, res = ctx.macproject(rel_eps=1e-10)
    scale = max(np.abs(O.valid(geom, ref["umac"][d][0], 0, 1, d)).max() for d in range(3))
    for d in range(3):
        got = download_like(ctx, geom, "UMAC_" + "XYZ"[d], ref["umac"][d], 1, 1)
        e = np.abs(O.valid(geom, got[0], 0, 1, d) - O.valid(geom, ref["umac"][d][0], 0, 1, d)).max() / scale
        assert e <= TOL_MAC, (d, e)
    ctx.close()
    print("ratio 1000: vcycles", ncyc, "res", res, "oracle cycles", ref["mac_cycles"])


@pytest.mark.parametrize("case", ["rt64", "mixed"])
@pytest.mark.parametrize("tile", [-1, 0, 1, 2, 3, 4])
def test_mg_fused_smoother(case, tile):
    """the fused smoother k_sweep3 (GSRB sweep + residual + restriction / prolongation / norm in one launch), every tile shape the
    launcher can pick (0: 32x32, 1: 64x16, 2: 32x16, 3: 64x14, 4: 32x24; -1: the production choice per launch kind), against the plain
    per-colour kernels and the oracle: same V-cycle, so phi and the projected velocity agree to the solver tolerance"""
    if case == "rt64":
        geom, P, st, dt = O.rt_state(64, dim=3, max_grid_size=64)
    else:
        # isotropic cells (dx = 1/64): point GSRB with piecewise-constant prolongation stalls on 2:1 anisotropic grids, in the oracle too
        geom, P, st, dt = O.random_state([48, 32, 64], dim=3, max_grid_size=64, phys_bc=[[IN, OUT], [NS, W], [PER, PER]], seed=11,
                                         prob_hi=[0.75, 0.5, 1.0])
    ref = O.stagewise(geom, P, st, dt, mac_rel_eps=1e-13)
    out = {}
    for mode in ("plain", "fused"):
        ctx = make_ctx(geom, P)
        if mode == "plain":
            ctx.mg_tune(1 << 30, -1)
        else:
            ctx.mg_tune(16, tile)
        upload_state(ctx, geom, P, st)
        ctx.mkvelforce("SOLD", 1.0)
        ctx.velpred(dt)
        try:
            ncyc, res = ctx.macproject(rel_eps=1e-10)
        except Exception as e:
            raise AssertionError("macproject failed in mode %s: %s" % (mode, e))
        assert res <= 1e-10, (mode, res)
        um = [download_like(ctx, geom, "UMAC_" + "XYZ"[d], ref["umac"][d], 1, 1) for d in range(3)]
        out[mode] = (ncyc, res, um, ctx.launch_count())
        ctx.close()
    scale = max(np.abs(O.valid(geom, ref["umac"][d][0], 0, 1, d)).max() for d in range(3))
    for d in range(3):
        a = O.valid(geom, out["fused"][2][d][0], 0, 1, d); b = O.valid(geom, out["plain"][2][d][0], 0, 1, d)
        r = O.valid(geom, ref["umac"][d][0], 0, 1, d)
        assert np.abs(a - b).max() / scale <= TOL_MAC, (d, np.abs(a - b).max() / scale)
        assert np.abs(a - r).max() / scale <= TOL_MAC
    # same algorithm => same cycle count (+-1 for round-off at the stopping test); far fewer launches
    assert abs(out["fused"][0] - out["plain"][0]) <= 1, (out["fused"][0], out["plain"][0])
    assert out["fused"][3] < out["plain"][3]
    print(case, tile, "cycles fused/plain", out["fused"][0], out["plain"][0], "launches", out["fused"][3], out["plain"][3])
