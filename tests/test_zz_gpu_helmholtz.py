"""
GPU parity (-m gpu) of SURVEY 8(f) row 2: vdn_visc_solve / vdn_diff_scalar_solve (viscsolve.f90:19,310) against the CPU oracle (itself
checked against a sparse direct solve, tests/test_helmholtz_cpu.py).  ** parity unpinned ** against F_MG like the MAC solve; both sides
converge to a relative RESIDUAL of 1e-12 (viscsolve.f90:88); the operators here have condition numbers of a few hundred, so the bar on the
SOLUTIONS is 1e-9.  (File name: runs after the parity tests of the path.)
"""
import numpy as np
import pytest

from oracle import oracle as O
from util import make_ctx, upload_state, relerr, download_like
from test_helmholtz_cpu import helm_state, CASES

pytestmark = pytest.mark.gpu
TOL = 1e-9
W, NS, IN, OUT, PER = O.SLIP_WALL, O.NO_SLIP_WALL, O.INLET, O.OUTLET, O.PERIODIC


def _ctx_with(geom, P, st, rho, lapu, mac_rhs, unew=None, snew=None):
    ctx = make_ctx(geom, P)
    upload_state(ctx, geom, P, st)
    ctx.upload_mf("RHOHALF", rho, 1, 1)
    ctx.upload_mf("LAPU", lapu, 0, geom.dim)
    ctx.upload_mf("MAC_RHS", mac_rhs, 1, 1)
    if unew is not None:
        ctx.upload_mf("UNEW", unew, 3, geom.dim)
    if snew is not None:
        ctx.upload_mf("SNEW", snew, 3, P.nscal)
    return ctx


@pytest.mark.parametrize("name", sorted(CASES) + ["3d_32_rtbc"])
@pytest.mark.parametrize("diffusion_type", [1, 2])
def test_visc_solve_matches_oracle(name, diffusion_type):
    if name == "3d_32_rtbc":
        n, dim, mgs, bc, seed = [32, 32, 32], 3, 16, [[PER, PER], [PER, PER], [NS, NS]], 35
    else:
        n, dim, mgs, bc, seed = CASES[name]
    geom, P, st, rho, lapu, mac_rhs = helm_state(n, dim, mgs, bc, seed)
    mu = 0.37 * min(geom.dx[:dim]) ** 2 * 40.0
    unew = [a.copy(order="F") for a in st["uold"]]
    ref = [a.copy(order="F") for a in unew]
    cyc_o, res_o = O.visc_solve(geom, P, ref, lapu, rho, mac_rhs, mu, diffusion_type)
    ctx = _ctx_with(geom, P, st, rho, lapu, mac_rhs, unew=unew)
    cyc, res = ctx.visc_solve(mu, diffusion_type)
    got = download_like(ctx, geom, "UNEW", ref, 3, dim)
    ctx.close()
    assert res <= 1e-12 and res_o <= 1e-12
    assert relerr(geom, got, ref, 3, full=False) <= TOL, (name, cyc, cyc_o)
    assert relerr(geom, got, ref, 3, full=True) <= TOL          # ghost cells refilled as viscsolve.f90:105 does


def test_diff_scalar_solve_matches_oracle():
    n, dim, mgs, bc, seed = CASES["3d_walls_inflow"]
    geom, P, st, rho, lapu, mac_rhs = helm_state(n, dim, mgs, bc, seed)
    mu = 0.5 * min(geom.dx[:dim]) ** 2 * 30.0
    snew = [a.copy(order="F") for a in st["sold"]]
    ref = [a.copy(order="F") for a in snew]
    laps = [np.zeros(a.shape[:3] + (P.nscal,), order="F") for a in lapu]
    O.diff_scalar_solve(geom, P, ref, laps, mu, 1, 2)
    ctx = _ctx_with(geom, P, st, rho, lapu, mac_rhs, snew=snew)
    cyc, res = ctx.diff_scalar_solve(mu, 1, 2)
    got = download_like(ctx, geom, "SNEW", ref, 3, P.nscal)
    ctx.close()
    assert res <= 1e-12
    assert relerr(geom, got, ref, 3, full=False) <= TOL


def test_refusals():
    import varden_b200 as V
    n, dim, mgs, bc, seed = CASES["3d_noslip_8box"]
    geom, P, st, rho, lapu, mac_rhs = helm_state(n, dim, mgs, bc, seed)
    ctx = make_ctx(geom, P)
    upload_state(ctx, geom, P, st)
    with pytest.raises(V.VdnError):
        ctx.visc_solve(1e-3, 1)                 # Crank-Nicolson without LAPU
    with pytest.raises(V.VdnError):
        ctx.diff_scalar_solve(1e-3, 0, 1)       # no LAPS field yet
    ctx.close()
