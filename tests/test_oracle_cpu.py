"""CPU tests of the oracle itself (the checker): invariants the reference arithmetic must satisfy (SURVEY 8(c) pins 2-4)
and the multigrid restatement against a sparse direct solve of the same discrete operator."""
import numpy as np
import pytest

from oracle import oracle as O

W, NS, IN, OUT, PER = O.SLIP_WALL, O.NO_SLIP_WALL, O.INLET, O.OUTLET, O.PERIODIC


def gather(geom, mf, ng, nc, face_dir=-1):
    n = geom.n_cell
    G = np.zeros((n[0], n[1], n[2], nc))
    for ib, (lo, hi) in enumerate(geom.boxes):
        v = O.valid(geom, mf[ib], ib, ng, face_dir)
        G[lo[0]:hi[0] + 1, lo[1]:hi[1] + 1, lo[2]:hi[2] + 1] = v[:hi[0] - lo[0] + 1, :hi[1] - lo[1] + 1, :hi[2] - lo[2] + 1]
    return G


@pytest.mark.parametrize("dim", [2, 3])
def test_constant_state_is_preserved(dim):
    """uniform u, rho with zero forcing: slopes vanish, every Riemann select returns the constant, update is exact"""
    geom, P, st, dt = O.rt_state(16, dim=dim, max_grid_size=8, phys_bc=[[PER, PER]] * dim)
    for ib in range(geom.nboxes):
        st["uold"][ib][...] = 0.0
        st["uold"][ib][..., 0] = 0.3
        st["uold"][ib][..., 1] = -0.2
        st["sold"][ib][..., 0] = 1.7
        st["sold"][ib][..., 1] = 0.4
        st["ext_vel_force"][ib][...] = 0.0
    out = O.advance(geom, P, st, dt)
    for ib in range(geom.nboxes):
        assert np.array_equal(out["snew"][ib], st["sold"][ib])
        assert np.allclose(out["unew"][ib], st["uold"][ib], rtol=0, atol=1e-14)


@pytest.mark.parametrize("dim", [2, 3])
def test_mass_is_conserved_with_periodic_bc(dim):
    """the conservative density update telescopes (update.f90:250-253): sum(rho) constant to round-off"""
    geom, P, st, dt = O.random_state(16, dim=dim, max_grid_size=8, phys_bc=[[PER, PER]] * dim, seed=11)
    out = O.advance(geom, P, st, dt)
    m0 = gather(geom, st["sold"], 3, 2)[..., 0].sum()
    m1 = gather(geom, out["snew"], 3, 2)[..., 0].sum()
    assert abs(m1 - m0) <= 1e-12 * abs(m0)


@pytest.mark.parametrize("dim,bc", [(3, [[PER, PER], [PER, PER], [NS, NS]]), (3, [[IN, OUT], [W, W], [PER, PER]]), (2, [[NS, NS], [W, W]])])
def test_box_decomposition_does_not_change_the_answer(dim, bc):
    """same domain as 1 box vs 8 (4) boxes: identical up to the per-box eps (never active for these states)"""
    a = O.random_state(16, dim=dim, max_grid_size=16, phys_bc=bc, seed=5)
    b = O.random_state(16, dim=dim, max_grid_size=8, phys_bc=bc, seed=5)
    oa = O.advance(a[0], a[1], a[2], a[3], mac_rel_eps=1e-12)
    ob = O.advance(b[0], b[1], b[2], b[3], mac_rel_eps=1e-12)
    for k, ng, nc in (("unew", 3, dim), ("snew", 3, 2)):
        assert np.array_equal(gather(a[0], oa[k], ng, nc), gather(b[0], ob[k], ng, nc)), k


def test_projection_removes_divergence():
    geom, P, st, dt = O.random_state(16, dim=3, max_grid_size=8, phys_bc=[[W, W], [PER, PER], [NS, NS]], seed=3)
    o = O.stagewise(geom, P, st, dt, mac_rel_eps=1e-12)
    um = [gather(geom, o["umac"][d], 1, 1, d) for d in range(3)]
    # gather() keeps only the lo faces of each box; rebuild the divergence from per-box arrays instead
    worst = 0.0
    for ib, (lo, hi) in enumerate(geom.boxes):
        u, v, w = (O.valid(geom, o["umac"][d][ib], ib, 1, d)[..., 0] for d in range(3))
        div = (u[1:, :, :] - u[:-1, :, :]) / geom.dx[0] + (v[:, 1:, :] - v[:, :-1, :]) / geom.dx[1] + (w[:, :, 1:] - w[:, :, :-1]) / geom.dx[2]
        worst = max(worst, np.abs(div).max())
    pre = 0.0
    for ib in range(geom.nboxes):
        u, v, w = (O.valid(geom, o["umac_pred"][d][ib], ib, 1, d)[..., 0] for d in range(3))
        div = (u[1:, :, :] - u[:-1, :, :]) / geom.dx[0] + (v[:, 1:, :] - v[:, :-1, :]) / geom.dx[1] + (w[:, :, 1:] - w[:, :, :-1]) / geom.dx[2]
        pre = max(pre, np.abs(div).max())
    assert worst <= 1e-11 * pre


@pytest.mark.parametrize("bc", [[[PER, PER], [PER, PER], [NS, NS]], [[IN, OUT], [W, W], [PER, PER]], [[OUT, OUT], [NS, W], [W, OUT]]])
def test_multigrid_matches_sparse_direct_solve(bc):
    """orc_mg_solve vs scipy.sparse direct solution of the identical stencil (mean removed when singular)"""
    import ctypes as C
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    rng = np.random.default_rng(0)
    n = [16, 8, 12]
    h = [1.0 / 16, 1.0 / 8, 1.0 / 12]
    ell = np.zeros((3, 2), dtype=np.int32)
    for d in range(3):
        for s in range(2):
            p = bc[d][s]
            ell[d, s] = -1 if p == PER else (1 if p == OUT else 2)
    singular = not (ell == 1).any()
    rho = np.exp(rng.uniform(-1.5, 1.5, size=[m + 2 for m in n]))
    for d in range(3):          # periodic images of rho
        if ell[d, 0] == -1:
            sl = [slice(None)] * 3
            lo, hi = list(sl), list(sl)
            lo[d], hi[d] = 0, -2
            rho[tuple(lo)] = rho[tuple(hi)]
            lo[d], hi[d] = -1, 1
            rho[tuple(lo)] = rho[tuple(hi)]
    c = rho[1:-1, 1:-1, 1:-1]
    bx = np.asfortranarray(2.0 / (rho[1:, 1:-1, 1:-1] + rho[:-1, 1:-1, 1:-1]))
    by = np.asfortranarray(2.0 / (rho[1:-1, 1:, 1:-1] + rho[1:-1, :-1, 1:-1]))
    bz = np.asfortranarray(2.0 / (rho[1:-1, 1:-1, 1:] + rho[1:-1, 1:-1, :-1]))
    beta = [bx, by, bz]
    rh = rng.standard_normal(n)
    if singular:
        rh -= rh.mean()
    rh = np.asfortranarray(rh)
    phi = np.zeros([m + 2 for m in n], order='F')
    res = C.c_double(0)
    f = O.lib().orc_mg_solve
    f.restype = C.c_int
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    na, ha = np.array(n, dtype=np.int32), np.array(h)
    cyc = f(3, ip(na), dp(ha), ip(ell), dp(rh), dp(bx), dp(by), dp(bz), dp(phi), C.c_double(1e-12), 200, 2, 2, C.c_double(1e-3), 0, C.byref(res))
    assert res.value <= 1e-12 and cyc < 200
    # assemble the same operator
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
    b = rh.ravel(order='F').copy()
    if singular:
        A = A + sp.csr_matrix(np.ones((1, N))).T @ sp.csr_matrix(np.ones((1, N))) / N      # pin the mean
    x = spla.spsolve(A.tocsc(), b)
    got = phi[1:-1, 1:-1, 1:-1].ravel(order='F')
    if singular:
        x -= x.mean(); got = got - got.mean()
    assert np.abs(got - x).max() <= 1e-9 * np.abs(x).max()
