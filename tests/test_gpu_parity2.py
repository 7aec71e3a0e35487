"""
GPU parity tests, second set (-m gpu): the holes the round-1 review named.
  * one step at 128^3 and 256^3 with the PRODUCTION kernel selection (fused smoother with its default tile shapes, marching Godunov kernels
    with their production tile) against the CPU oracle, including phi (mean removed: the RT operator is singular);
  * the device multigrid against a sparse DIRECT solve of the same discrete operator (the strongest pin the solver can have while F_MG,
    the reference's third-party multigrid, is absent: parity of the MAC solve stays "unpinned" against F_MG itself);
  * BASELINE config 1: the reference's own CPU deck (exec/test/inputs_2d-regt, max_levs = 1: 2-D 64^2 bubble, 4 boxes, no-slip walls).
Tolerances (BASELINE.json north_star): edge states / updated fields 1e-12 on identical inputs; projected velocity and phi 10 x the MAC
tolerance (1e-9); whole step with both solves converged to 1e-13: 1e-10.
"""
import numpy as np
import pytest

from oracle import oracle as O
from util import make_ctx, upload_state, relerr, download_like

pytestmark = pytest.mark.gpu
TOL_MAC = 1e-9
W, NS, IN, OUT, PER = O.SLIP_WALL, O.NO_SLIP_WALL, O.INLET, O.OUTLET, O.PERIODIC


def _demean(geom, mf, ng):
    v = [O.valid(geom, a, ib, ng)[..., 0] for ib, a in enumerate(mf)]
    m = sum(x.sum() for x in v) / sum(x.size for x in v)
    return [x - m for x in v]


@pytest.mark.parametrize("n", [128, 256])
def test_one_step_production_kernels(n):
    """default settings: nothing forced -- at these sizes the launcher picks the fused smoother (k_sweep3, tile shapes 4 / 2 / 4) on the fine
    levels and the marching Godunov kernels; MAC tolerance 1e-12 on both sides for the phi / umac comparison, then the whole step"""
    geom, P, st, dt = O.rt_state(n, dim=3, max_grid_size=n)
    ref = O.stagewise(geom, P, st, dt, mac_rel_eps=1e-12)
    ctx = make_ctx(geom, P)
    upload_state(ctx, geom, P, st)
    ctx.mkvelforce("SOLD", 1.0)
    ctx.velpred(dt)
    for d in range(3):
        got = download_like(ctx, geom, "UMAC_" + "XYZ"[d], ref["umac_pred"][d], 1, 1)
        assert relerr(geom, got, ref["umac_pred"][d], 1, d) <= 1e-12
    ncyc, res = ctx.macproject(rel_eps=1e-12)
    assert res <= 1e-12
    scale = max(np.abs(O.valid(geom, ref["umac"][d][0], 0, 1, d)).max() for d in range(3))
    for d in range(3):
        got = download_like(ctx, geom, "UMAC_" + "XYZ"[d], ref["umac"][d], 1, 1)
        e = np.abs(O.valid(geom, got[0], 0, 1, d) - O.valid(geom, ref["umac"][d][0], 0, 1, d)).max() / scale
        assert e <= TOL_MAC, ("umac", d, e)
    phi = download_like(ctx, geom, "PHI", ref["phi"], 1, 1)
    a, b = _demean(geom, phi, 1), _demean(geom, ref["phi"], 1)
    e_phi = max(np.abs(x - y).max() for x, y in zip(a, b)) / max(np.abs(y).max() for y in b)
    assert e_phi <= TOL_MAC, ("phi", e_phi)
    # the whole step, both solves converged as far as FP64 allows at this size (SURVEY Q10 mode b): 1e-13 at 128^3; at 256^3 the relative
    # residual stalls above 1e-13 (round-off of a 1.7e7-cell operator), so 1e-12 there
    eps = 1e-13 if n <= 128 else 1e-12
    ref2 = O.advance(geom, P, st, dt, mac_rel_eps=eps)
    upload_state(ctx, geom, P, st)
    ctx.advance(dt, mac_rel_eps=eps)
    e_s = relerr(geom, download_like(ctx, geom, "SNEW", ref2["snew"], 3, P.nscal), ref2["snew"], 3)
    e_u = relerr(geom, download_like(ctx, geom, "UNEW", ref2["unew"], 3, 3), ref2["unew"], 3)
    ctx.close()
    print("n=%d: V-cycles %d, phi err %.2e, snew %.2e, unew %.2e" % (n, ncyc, e_phi, e_s, e_u))
    assert e_s <= 1e-10 and e_u <= 1e-10


@pytest.mark.parametrize("fuse_min", [1 << 30, 16])
@pytest.mark.parametrize("bc", [[[PER, PER], [PER, PER], [NS, NS]], [[IN, OUT], [W, W], [PER, PER]], [[OUT, OUT], [NS, W], [W, OUT]]],
                         ids=["per_per_wall", "inout_wall_per", "out_wall_mixed"])
def test_device_multigrid_vs_sparse_direct_solve(bc, fuse_min):
    """vdn_mac_solve (plain kernels / fused smoother) against scipy's sparse direct solution of the identical stencil: Neumann faces drop
    the coefficient, Dirichlet (OUTLET) faces use the stencil_order = 2 one-sided form, the singular case is compared mean-free"""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    rng = np.random.default_rng(3)
    n = [32, 16, 16]
    geom = O.Geom(3, n, bc, prob_hi=[1.0, 0.5, 0.5], max_grid_size=32)       # dx = 1/32 in every direction
    P = O.Params(dim=3, nscal=2)
    h = geom.dx
    ell = np.zeros((3, 2), dtype=np.int32)
    for d in range(3):
        for s in range(2):
            ell[d, s] = -1 if bc[d][s] == PER else (1 if bc[d][s] == OUT else 2)
    singular = not (ell == 1).any()
    rho = np.exp(rng.uniform(-1.5, 1.5, size=[m + 2 for m in n]))
    for d in range(3):
        if ell[d, 0] == -1:
            lo, hi = [slice(None)] * 3, [slice(None)] * 3
            lo[d], hi[d] = 0, -2
            rho[tuple(lo)] = rho[tuple(hi)]
            lo[d], hi[d] = -1, 1
            rho[tuple(lo)] = rho[tuple(hi)]
    beta = [np.asfortranarray(2.0 / (rho[1:, 1:-1, 1:-1] + rho[:-1, 1:-1, 1:-1])),
            np.asfortranarray(2.0 / (rho[1:-1, 1:, 1:-1] + rho[1:-1, :-1, 1:-1])),
            np.asfortranarray(2.0 / (rho[1:-1, 1:-1, 1:] + rho[1:-1, 1:-1, :-1]))]
    rh = rng.standard_normal(n)
    if singular:
        rh -= rh.mean()
    ctx = make_ctx(geom, P)
    ctx.mg_tune(fuse_min, -1)
    ctx.upload_mf("RH", [np.asfortranarray(rh[..., None])], 0, 1)
    for d in range(3):
        ctx.upload_mf("BETA_" + "XYZ"[d], [np.asfortranarray(beta[d][..., None])], 0, 1)
    ctx.setval("PHI", 0.0)
    ncyc, res = ctx.mac_solve(rel_eps=1e-12)
    assert res <= 1e-12
    out = [np.full([m + 2 for m in n] + [1], np.nan, order="F")]
    ctx.download_mf("PHI", out, 1, 1)
    ctx.close()
    got = out[0][1:-1, 1:-1, 1:-1, 0].ravel(order="F")
    N = n[0] * n[1] * n[2]
    idx = lambda i, j, k: i + n[0] * (j + n[1] * k)
    rows, cols, vals = [], [], []
    for k in range(n[2]):
        for j in range(n[1]):
            for i in range(n[0]):
                ix, me, diag = (i, j, k), idx(i, j, k), 0.0
                for d in range(3):
                    h2 = 1.0 / h[d] ** 2
                    for side, off in ((0, -1), (1, 1)):
                        fidx = list(ix)
                        if side == 1:
                            fidx[d] += 1
                        b = beta[d][tuple(fidx)]
                        at_b = (ix[d] == 0) if side == 0 else (ix[d] == n[d] - 1)
                        nb = list(ix)
                        nb[d] += off
                        if at_b and ell[d, side] == 2:
                            continue
                        if at_b and ell[d, side] == 1:
                            inner = list(ix)
                            inner[d] -= off
                            diag += 3.0 * b * h2
                            rows.append(me); cols.append(idx(*inner)); vals.append(-b * h2 / 3.0)
                            continue
                        nb[d] %= n[d]
                        diag += b * h2
                        rows.append(me); cols.append(idx(*nb)); vals.append(-b * h2)
                rows.append(me); cols.append(me); vals.append(diag)
    A = sp.csr_matrix((vals, (rows, cols)), shape=(N, N))
    b = rh.ravel(order="F").copy()
    if singular:
        A = A + sp.csr_matrix(np.ones((1, N))).T @ sp.csr_matrix(np.ones((1, N))) / N      # pin the mean
    x = spla.spsolve(A.tocsc(), b)
    if singular:
        x -= x.mean(); got = got - got.mean()
    err = np.abs(got - x).max() / np.abs(x).max()
    print("direct solve: V-cycles %d, |phi - phi_direct| / |phi| = %.2e" % (ncyc, err))
    assert err <= TOL_MAC


def test_config1_bubble_2d():
    """BASELINE config 1: exec/test/inputs_2d-regt with max_levs = 1 (64^2, four 32^2 boxes, prob_type 1 bubble, no-slip walls), inviscid path"""
    geom, P, st, dt = O.bubble_state(64, max_grid_size=32)
    assert geom.nboxes == 4
    state = {k: [a.copy(order="F") for a in v] for k, v in st.items()}
    ctx = make_ctx(geom, P)
    worst = 0.0
    for step in range(5):
        ref = O.advance(geom, P, state, dt, mac_rel_eps=1e-13)
        upload_state(ctx, geom, P, state)
        ncyc, res = ctx.advance(dt, mac_rel_eps=1e-13)
        e_s = relerr(geom, download_like(ctx, geom, "SNEW", ref["snew"], 3, P.nscal), ref["snew"], 3, full=True)
        e_u = relerr(geom, download_like(ctx, geom, "UNEW", ref["unew"], 3, 2), ref["unew"], 3, full=True)
        e_r = relerr(geom, download_like(ctx, geom, "RHOHALF", ref["rhohalf"], 1, 1), ref["rhohalf"], 1, full=True)
        worst = max(worst, e_s, e_u, e_r)
        print("bubble step %d: V-cycles %d, snew %.2e unew %.2e rhohalf %.2e" % (step, ncyc, e_s, e_u, e_r))
        # next step from the ORACLE's state (u* is not projected here: hgproject stays the reference's; the path test only needs a
        # sequence of different, physically shaped inputs)
        state = dict(state, uold=ref["unew"], sold=ref["snew"])
    ctx.close()
    assert worst <= 1e-10


@pytest.mark.parametrize("case", ["rt3d_8box", "rand3d_inout", "rand2d_4box", "tiny_velocities"])
def test_estdt_on_device_bit_exact(case):
    """estdt (estdt.f90:15-181) on the resident fields equals the oracle's dt (itself bit-identical to the reference's own
    estdt_2d / estdt_3d, tests/test_ref_pin.py) bit for bit -- the maxima are exact, the limits are computed in the reference's order;
    then the driver's uold <- unew copy (varden.f90:321-324) on the device."""
    if case == "rt3d_8box":
        geom, P, st, dt = O.rt_state(32, dim=3, max_grid_size=16)
    elif case == "rand3d_inout":
        geom, P, st, dt = O.random_state([24, 16, 20], dim=3, max_grid_size=12, phys_bc=[[IN, OUT], [PER, PER], [NS, W]], seed=21)
    elif case == "rand2d_4box":
        geom, P, st, dt = O.random_state([32, 32], dim=2, max_grid_size=16, phys_bc=[[NS, NS], [IN, OUT]], seed=22)
    else:
        geom, P, st, dt = O.random_state([16, 16, 16], dim=3, max_grid_size=16, phys_bc=[[W, W]] * 3, seed=23)
        for k in ("uold", "gp", "ext_vel_force"):
            st[k] = [a * 1e-12 for a in st[k]]              # everything below the reference's eps: the min(dx) fallback
    ctx = make_ctx(geom, P)
    upload_state(ctx, geom, P, st)
    for dtold in (-1.0, 1e-4):
        want = O.estdt(geom, st["uold"], 3, st["sold"], 3, st["gp"], 1, st["ext_vel_force"], 1, dtold=dtold)
        got = ctx.estdt(dtold=dtold)
        assert got == want, (case, dtold, got, want)
    ctx.advance(dt)
    ctx.field_copy("UOLD", "UNEW")
    unew = download_like(ctx, geom, "UNEW", st["uold"], 3, geom.dim)
    uold = download_like(ctx, geom, "UOLD", st["uold"], 3, geom.dim)
    for a, b in zip(unew, uold):
        assert np.array_equal(a, b)
    ctx.close()
