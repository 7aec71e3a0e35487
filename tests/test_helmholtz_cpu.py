"""
SURVEY 8(f) row 2 -- visc_solve / diff_scalar_solve (viscsolve.f90): the oracle's Helmholtz form of the multigrid, ** parity unpinned **
like the MAC solve (F_MG is absent from the reference tree), checked against a sparse direct solve of the same discrete system
assembled independently here:  (alpha - mu div grad) phi = rh  with the stencil_order-2 boundary stencil; an EXT_DIR ghost cell holds
the boundary value phi_b and contributes 8/3 mu phi_b / h^2 to the right-hand side.
"""
import numpy as np
import pytest

from oracle import oracle as O

W, NS, IN, OUT, PER, SYM = O.SLIP_WALL, O.NO_SLIP_WALL, O.INLET, O.OUTLET, O.PERIODIC, O.SYMMETRY


def helm_state(n, dim, mgs, bc, seed, prob_hi=None):
    """a random state plus what the viscous solve reads: rhohalf (ng 1), lapu (ng 0), mac_rhs (ng 1, ghost cells filled)"""
    geom, P, st, dt = O.random_state(n, dim=dim, max_grid_size=mgs, phys_bc=bc, seed=seed, prob_hi=prob_hi)
    rng = np.random.default_rng(100 + seed)
    rho = O.mf_alloc(geom, 1, 1)
    for ib in range(geom.nboxes):
        sl = tuple(slice(2, -2) if d < dim else slice(None) for d in range(3))
        rho[ib][..., 0] = st["sold"][ib][sl + (0,)]
    lapu = [np.asfortranarray(rng.standard_normal(a.shape)) for a in O.mf_alloc(geom, 0, dim)]
    mac_rhs = O.mf_alloc(geom, 1, 1)
    N = [geom.n_cell[d] for d in range(3)]
    G = rng.standard_normal([N[d] + (2 if d < dim else 0) for d in range(3)])
    for ib, (lo, hi) in enumerate(geom.boxes):
        sl = tuple(slice(lo[d], hi[d] + 3) if d < dim else slice(None) for d in range(3))
        mac_rhs[ib][..., 0] = G[sl]
    return geom, P, st, rho, lapu, mac_rhs


CASES = {
    "3d_walls_inflow": ([16, 8, 8], 3, 8, [[IN, OUT], [NS, W], [PER, PER]], 31),
    "3d_noslip_8box": ([8, 8, 8], 3, 4, [[NS, NS], [NS, NS], [NS, NS]], 32),
    "2d_slip_sym": ([16, 8], 2, 8, [[W, SYM], [NS, IN]], 33),
}


def direct_solve(geom, u_mf, comp, comp_is_vel, rho, lapu, mac_rhs, mu, diffusion_type):
    """independent assembly: loops over cells, scipy sparse LU"""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    dim = geom.dim
    n = [geom.n_cell[d] for d in range(3)]
    ug = O._gather(geom, u_mf, 3, comp, grow=1)
    al = O._gather(geom, rho, 1, 0) if comp_is_vel else np.ones(n)
    lap = O._gather(geom, lapu, 0, comp)
    mg = O._gather(geom, mac_rhs, 1, 0, grow=1) if comp_is_vel else None
    ell = O.helm_ell_bc(geom, comp_is_vel, comp)
    visc_mu_dt = 2.0 * mu if diffusion_type == 1 else mu
    g = lambda ix: tuple(ix[d] + 1 if d < dim else 0 for d in range(3))
    idx = lambda ix: ix[0] + n[0] * (ix[1] + n[1] * ix[2])
    rows, cols, vals = [], [], []
    b = np.zeros(n[0] * n[1] * n[2])
    for k in range(n[2]):
        for j in range(n[1]):
            for i in range(n[0]):
                ix = (i, j, k)
                me = idx(ix)
                u = ug[g(ix)]
                rh = u * al[ix] if comp_is_vel else u
                if diffusion_type == 1:
                    rh += mu * lap[ix]
                if comp_is_vel:
                    p, m = list(ix), list(ix)
                    p[comp] += 1; m[comp] -= 1
                    rh += (1.0 / 3.0) * visc_mu_dt * (mg[g(p)] - mg[g(m)]) / geom.dx[comp]
                diag = al[ix]
                for d in range(dim):
                    h2 = 1.0 / geom.dx[d] ** 2
                    for side, off in ((0, -1), (1, 1)):
                        at_b = ix[d] == 0 if side == 0 else ix[d] == n[d] - 1
                        nb = list(ix); nb[d] += off
                        if at_b and ell[d, side] == 2:
                            continue
                        if at_b and ell[d, side] == 1:
                            inner = list(ix); inner[d] -= off
                            diag += 3.0 * mu * h2
                            rows.append(me); cols.append(idx(inner)); vals.append(-mu * h2 / 3.0)
                            rh += (8.0 / 3.0) * mu * h2 * ug[g(nb)]
                            continue
                        nb[d] %= n[d]
                        diag += mu * h2
                        rows.append(me); cols.append(idx(nb)); vals.append(-mu * h2)
                rows.append(me); cols.append(me); vals.append(diag)
                b[me] = rh
    A = sp.csr_matrix((vals, (rows, cols)), shape=(b.size, b.size))
    return spla.spsolve(A.tocsc(), b).reshape(n, order="F")


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("diffusion_type", [1, 2])
def test_visc_solve_matches_sparse_direct_solve(name, diffusion_type):
    n, dim, mgs, bc, seed = CASES[name]
    geom, P, st, rho, lapu, mac_rhs = helm_state(n, dim, mgs, bc, seed)
    mu = 0.37 * min(geom.dx[:dim]) ** 2 * 40.0              # mu / h^2 ~ 15: the operator is far from the identity
    unew = [a.copy(order="F") for a in st["uold"]]
    want = [direct_solve(geom, unew, d, True, rho, lapu, mac_rhs, mu, diffusion_type) for d in range(dim)]
    cyc, res = O.visc_solve(geom, P, unew, lapu, rho, mac_rhs, mu, diffusion_type)
    assert res <= 1e-12 and cyc < 60 * dim
    for d in range(dim):
        got = O._gather(geom, unew, 3, d)
        assert np.abs(got - want[d]).max() <= 1e-10 * np.abs(want[d]).max(), (name, d)


def test_diff_scalar_solve_matches_sparse_direct_solve():
    n, dim, mgs, bc, seed = CASES["3d_walls_inflow"]
    geom, P, st, rho, lapu, mac_rhs = helm_state(n, dim, mgs, bc, seed)
    mu = 0.5 * min(geom.dx[:dim]) ** 2 * 30.0
    snew = [a.copy(order="F") for a in st["sold"]]
    laps = [np.zeros(a.shape[:3] + (P.nscal,), order="F") for a in lapu]
    want = direct_solve(geom, snew, 1, False, None, laps, None, mu, 2)
    cyc, res = O.diff_scalar_solve(geom, P, snew, laps, mu, 1, 2)
    assert res <= 1e-12
    got = O._gather(geom, snew, 3, 1)
    assert np.abs(got - want).max() <= 1e-10 * np.abs(want).max()


def test_mu_zero_is_the_identity():
    n, dim, mgs, bc, seed = CASES["3d_noslip_8box"]
    geom, P, st, rho, lapu, mac_rhs = helm_state(n, dim, mgs, bc, seed)
    unew = [a.copy(order="F") for a in st["uold"]]
    O.visc_solve(geom, P, unew, lapu, rho, mac_rhs, 0.0, 2)
    for a, b in zip(unew, st["uold"]):
        va, vb = O.valid(geom, a, 0, 3), O.valid(geom, b, 0, 3)
        assert np.abs(va - vb).max() <= 1e-13 * np.abs(vb).max()
