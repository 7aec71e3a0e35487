"""
CPU execution of the REAL k_sweep3 kernel source (varden_b200/csrc/vdn_mg_fused.cuh), in the PRODUCTION tile shapes, under tests/emu/cuda_emu.h
(one OS thread per CUDA thread, std::barrier for __syncthreads), checked against a plain numpy red-black
Gauss-Seidel / residual / restriction / prolongation of the same operator (mac_multigrid.f90:53-62 selects these).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")
M_GHOST, M_NEU, M_DIR, M_WRAP = 0, 1, 2, 3
PAD = 4          # MG_PAD in vdn_ctx.h


@pytest.fixture(scope="module")
def emu():
    dev = os.environ.get("VDN_FUSED_HEADER")                     # development: test a working copy of the kernel header
    so = os.path.join(EMU, "libemu_wave%s.so" % ("_dev" if dev else ""))
    src = [os.path.join(EMU, "emu_wave.cpp"), os.path.join(EMU, "cuda_emu.h"),
           dev or os.path.join(HERE, "..", "varden_b200", "csrc", "vdn_mg_fused.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["g++", "-O1", "-std=c++20", "-pthread", "-fPIC", "-shared", "-ffp-contract=off"] +
                              (['-DFUSED_HEADER="%s"' % dev] if dev else []) + [src[0], "-o", so])
    return C.CDLL(so)


def pad(n):
    return (n[2] + 2 * PAD, n[1] + 2 * PAD, n[0] + 2 * PAD)


def fill_wrap(p, n, mode):
    for d, ax in ((0, 2), (1, 1), (2, 0)):
        if mode[d][0] == M_WRAP:
            sl_lo = [slice(None)] * 3; sl_hi = [slice(None)] * 3; src_lo = [slice(None)] * 3; src_hi = [slice(None)] * 3
            sl_lo[ax] = PAD - 1; src_lo[ax] = PAD - 1 + n[d]
            sl_hi[ax] = PAD + n[d]; src_hi[ax] = PAD
            p[tuple(sl_lo)] = p[tuple(src_lo)]
            p[tuple(sl_hi)] = p[tuple(src_hi)]


def apply_A(phi, b, h2, mode, n):
    """A*phi and diag on valid cells; b[d][idx] = coefficient on the LOW d-face of cell idx (padded arrays, z,y,x order)."""
    p = phi.copy()
    fill_wrap(p, n, mode)
    V = (slice(PAD, n[2] + PAD), slice(PAD, n[1] + PAD), slice(PAD, n[0] + PAD))
    p0 = p[V]
    ax = np.zeros_like(p0); dg = np.zeros_like(p0)
    for d, axis in ((0, 2), (1, 1), (2, 0)):
        def sh(a, o):
            s = list(V); s[axis] = slice(PAD + o, n[d] + PAD + o); return a[tuple(s)]
        blo, bhi = sh(b[d], 0), sh(b[d], 1)
        pm, pp = sh(p, -1), sh(p, 1)
        idx = np.arange(n[d]).reshape([-1 if a == axis else 1 for a in range(3)])
        atlo, athi = idx == 0, idx == n[d] - 1
        lo_reg = blo * (p0 - pm) * h2[d]; hi_reg = bhi * (p0 - pp) * h2[d]
        lo_dir = blo * (3.0 * p0 - pp / 3.0) * h2[d]; hi_dir = bhi * (3.0 * p0 - pm / 3.0) * h2[d]
        mlo, mhi = mode[d]
        lo = np.where(atlo & (mlo == M_NEU), 0.0, np.where(atlo & (mlo == M_DIR), lo_dir, lo_reg))
        hi = np.where(athi & (mhi == M_NEU), 0.0, np.where(athi & (mhi == M_DIR), hi_dir, hi_reg))
        glo = np.where(atlo & (mlo == M_NEU), 0.0, np.where(atlo & (mlo == M_DIR), 3.0, 1.0)) * blo * h2[d]
        ghi = np.where(athi & (mhi == M_NEU), 0.0, np.where(athi & (mhi == M_DIR), 3.0, 1.0)) * bhi * h2[d]
        ax += lo + hi; dg += glo + ghi
    return ax, dg


def gsrb(phi, rhs, b, h2, mode, n, par0, sweeps):
    V = (slice(PAD, n[2] + PAD), slice(PAD, n[1] + PAD), slice(PAD, n[0] + PAD))
    k, j, i = np.meshgrid(np.arange(n[2]), np.arange(n[1]), np.arange(n[0]), indexing="ij")
    par = (i + j + k + par0) & 1
    phi = phi.copy()
    for _ in range(sweeps):
        for color in (0, 1):
            ax, dg = apply_A(phi, b, h2, mode, n)
            upd = phi[V] + (rhs[V] - ax) / dg
            phi[V] = np.where(par == color, upd, phi[V])
    return phi


CASES = [
    # n, cfg (tile shape: 0 32x32, 1 64x16, 2 32x16, 3 64x14, 4 32x24, 5 16x8), zchunk, mode, par0
    ((64, 32, 16), 2, 8, ((M_WRAP, M_WRAP), (M_WRAP, M_WRAP), (M_NEU, M_NEU)), 0),
    ((64, 32, 16), 5, 16, ((M_NEU, M_DIR), (M_DIR, M_NEU), (M_NEU, M_NEU)), 1),
    ((64, 16, 8), 1, 4, ((M_WRAP, M_WRAP), (M_NEU, M_NEU), (M_WRAP, M_WRAP)), 0),
    ((40, 24, 12), 5, 6, ((M_NEU, M_NEU), (M_WRAP, M_WRAP), (M_DIR, M_DIR)), 0),     # ragged tiles
    ((16, 16, 16), 5, 16, ((M_DIR, M_DIR), (M_DIR, M_NEU), (M_NEU, M_DIR)), 1),
    ((64, 48, 8), 4, 8, ((M_WRAP, M_WRAP), (M_NEU, M_DIR), (M_NEU, M_NEU)), 1),      # 32x24: the production tile of the level-0 sweeps
    ((64, 28, 8), 3, 4, ((M_DIR, M_NEU), (M_WRAP, M_WRAP), (M_WRAP, M_WRAP)), 0),    # 64x14
    ((32, 64, 8), 0, 8, ((M_NEU, M_NEU), (M_WRAP, M_WRAP), (M_DIR, M_NEU)), 0),      # 32x32
]


KERNELS = [("sweep3", 1, pre, post) for pre in (0, 1) for post in (0, 2, 3)]


@pytest.mark.parametrize("n,cfg,zchunk,mode,par0", CASES)
@pytest.mark.parametrize("kern,nsw,pre,post", KERNELS)
def test_wave_matches_plain_gsrb(emu, n, cfg, zchunk, mode, par0, kern, nsw, pre, post):
    rng = np.random.default_rng(1234 + n[0] + 7 * nsw + pre + 3 * post)
    shp = pad(n)
    cn = tuple(x // 2 for x in n)
    h2 = np.array([1.0 / 0.01 ** 2, 1.0 / 0.012 ** 2, 1.0 / 0.009 ** 2])
    b = [np.ascontiguousarray(0.5 + rng.random(shp)) for _ in range(3)]
    # periodic: the face at index n equals the face at index 0
    for d, axis in ((0, 2), (1, 1), (2, 0)):
        if mode[d][0] == M_WRAP:
            s_hi = [slice(None)] * 3; s_lo = [slice(None)] * 3
            s_hi[axis] = n[d] + PAD; s_lo[axis] = PAD
            b[d][tuple(s_hi)] = b[d][tuple(s_lo)]
    rhs = np.ascontiguousarray(rng.standard_normal(shp))
    phi = np.ascontiguousarray(rng.standard_normal(shp))
    cphi = np.ascontiguousarray(rng.standard_normal(pad(cn)))
    V = (slice(PAD, n[2] + PAD), slice(PAD, n[1] + PAD), slice(PAD, n[0] + PAD))
    CV = (slice(PAD, cn[2] + PAD), slice(PAD, cn[1] + PAD), slice(PAD, cn[0] + PAD))
    out = np.full(shp, np.nan)
    crhs = np.full(pad(cn), np.nan); czero = np.full(pad(cn), np.nan)
    nrm = np.zeros(1)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    fn = getattr(emu, "emu_" + kern)
    _, dg0 = apply_A(phi, b, h2, mode, n)                     # the kernel takes 1 / diagonal (k_diag_inv in vdn_mg.cu computes it per solve)
    dinv = np.zeros(shp); dinv[V] = np.where(dg0 != 0.0, 1.0 / np.where(dg0 != 0.0, dg0, 1.0), 0.0)
    rc = fn(nsw, pre, post, cfg, (C.c_int * 3)(*n), (C.c_int * 6)(*[m for d in mode for m in d]), par0, P(h2),
            P(rhs), P(b[0]), P(b[1]), P(b[2]), P(phi), P(out), P(cphi), P(crhs), P(czero), P(nrm), zchunk, PAD, P(dinv))
    assert rc == 0
    # reference
    start = phi.copy()
    if pre:
        start[V] += np.repeat(np.repeat(np.repeat(cphi[CV], 2, axis=0), 2, axis=1), 2, axis=2)
    ref = gsrb(start, rhs, b, h2, mode, n, par0, nsw)
    scale = np.abs(ref[V]).max()
    assert np.all(np.isfinite(out[V]))
    assert np.abs(out[V] - ref[V]).max() <= 1e-12 * scale
    if post:
        ax, _ = apply_A(ref, b, h2, mode, n)
        res = rhs[V] - ax
        rs = np.abs(res).max()
        if post == 3:
            assert abs(nrm[0] - rs) <= 1e-10 * rs
        else:
            cr = res.reshape(cn[2], 2, cn[1], 2, cn[0], 2).mean(axis=(1, 3, 5))
            assert np.abs(crhs[CV] - cr).max() <= 1e-10 * rs
            assert np.all(czero[CV] == 0.0)


GM_PER = ((M_WRAP, M_WRAP),) * 3
GM_MIX = ((M_NEU, M_DIR), (M_WRAP, M_WRAP), (M_NEU, M_NEU))        # inflow / outflow x, periodic y, walls z (the rand3d multi-GPU case)


@pytest.mark.parametrize("cfg,pre,post", [(1, 0, 0), (1, 1, 0), (4, 0, 3), (2, 0, 2)])
def test_wave_rank_ghost_layers_production_tiles(emu, cfg, pre, post):
    """the tile shapes the launcher picks on a 16^3 rank-local level (64x16 plain / prolongating sweeps, 32x24 sweep + norm, 32x16 sweep +
    residual + restriction): one ragged CTA whose tile is larger than the level, ghost layers on every shared face, mixed physical boundaries,
    2 x 2 x 2 ranks in peer-memory mode -- the configuration of the 8-GPU rand3d parity case"""
    test_wave_rank_ghost_layers(emu, pre, post, (0, 1, 2), GM_MIX, (32, 32, 32), cfg, 1)


@pytest.mark.parametrize("p2p", [0, 1])
@pytest.mark.parametrize("cfg", [5, 2])
@pytest.mark.parametrize("pre,post", [(0, 0), (1, 0), (0, 2), (1, 3)])
@pytest.mark.parametrize("split,gmode,N", [((0,), GM_PER, (32, 32, 16)), ((1, 2), GM_PER, (32, 32, 32)), ((0, 1, 2), GM_PER, (32, 32, 32)),
                                           ((0, 1, 2), GM_MIX, (32, 32, 32)), ((0, 2), GM_MIX, (32, 16, 32))])
def test_wave_rank_ghost_layers(emu, pre, post, split, gmode, N, cfg, p2p, kern="sweep3"):
    """a level split across ranks: the kernel relaxes the neighbour ranks' cells redundantly (they sit in the rank's MG_PAD ghost layers,
    M_GHOST); the block of every 'rank' must equal the same block of the whole-domain sweep.  p2p = 1: the peer-memory mode -- every rank's
    out / coarse rhs / coarse phi arrays live in one buffer (the symmetric heap) and the kernel also stores what it writes near a shared face into
    the ghost layers of the neighbours' arrays: after all ranks have run, the PUSH_DEPTH = 3 ghost layers of those arrays (faces, edges, corners)
    must hold the whole-domain result as well -- the next launch reads them instead of waiting for an exchange"""
    rng = np.random.default_rng(77 + pre + 5 * post + len(split))
    n = tuple(N[d] // 2 if d in split else N[d] for d in range(3))
    cN = tuple(x // 2 for x in N); cn = tuple(x // 2 for x in n)
    per = [gmode[d][0] == M_WRAP for d in range(3)]
    h2 = np.array([1.0e4, 0.8e4, 1.3e4])
    shpN = pad(N)
    b = [np.ascontiguousarray(0.5 + rng.random(shpN)) for _ in range(3)]
    rhs = np.ascontiguousarray(rng.standard_normal(shpN)); phi = np.ascontiguousarray(rng.standard_normal(shpN))
    cphi = np.ascontiguousarray(rng.standard_normal(pad(cN)))

    def periodic_fill(a, nn):             # all PAD ghost layers of a whole-domain array along the periodic directions
        for ax, d in ((2, 0), (1, 1), (0, 2)):
            if per[d]:
                idx = (np.arange(-PAD, nn[d] + PAD) % nn[d]) + PAD
                a[...] = np.take(a, idx, axis=ax)
    for a in b + [rhs, phi]:
        periodic_fill(a, N)
    periodic_fill(cphi, cN)
    VN = (slice(PAD, N[2] + PAD), slice(PAD, N[1] + PAD), slice(PAD, N[0] + PAD))
    CVN = (slice(PAD, cN[2] + PAD), slice(PAD, cN[1] + PAD), slice(PAD, cN[0] + PAD))
    start = phi.copy()
    if pre:
        start[VN] += np.repeat(np.repeat(np.repeat(cphi[CVN], 2, axis=0), 2, axis=1), 2, axis=2)
    ref = gsrb(start, rhs, b, h2, gmode, N, 0, 1)
    _, dgN = apply_A(phi, b, h2, gmode, N)
    dinvN = np.zeros(shpN); dinvN[VN] = 1.0 / dgN
    periodic_fill(dinvN, N)
    ax, _ = apply_A(ref, b, h2, gmode, N)
    res = rhs[VN] - ax
    cres = np.zeros(pad(cN)); cres[CVN] = res.reshape(cN[2], 2, cN[1], 2, cN[0], 2).mean(axis=(1, 3, 5))
    refg = ref.copy()
    periodic_fill(refg, N); periodic_fill(cres, cN)
    Pp = lambda a: a.ctypes.data_as(C.c_void_p)
    pgrid = [2 if d in split else 1 for d in range(3)]
    corners = list(np.ndindex(*pgrid))
    def rank_mode(corner):
        return tuple(((M_GHOST if (corner[d] > 0 or per[d]) else gmode[d][0]), (M_GHOST if (corner[d] < pgrid[d] - 1 or per[d]) else gmode[d][1]))
                     if d in split else gmode[d] for d in range(3))
    def cut(a, nn, oo):                                # the block with its ghost layers, as the rank stores it
        return np.ascontiguousarray(a[oo[2]:oo[2] + nn[2] + 2 * PAD, oo[1]:oo[1] + nn[1] + 2 * PAD, oo[0]:oo[0] + nn[0] + 2 * PAD])
    # "symmetric heap" of every rank: out | crhs | czero
    szf, szc = int(np.prod(pad(n))), int(np.prod(pad(cn)))
    heap = {c_: np.full(szf + 2 * szc, np.nan) for c_ in corners}
    view = lambda c_: (heap[c_][:szf].reshape(pad(n)), heap[c_][szf:szf + szc].reshape(pad(cn)), heap[c_][szf + szc:].reshape(pad(cn)))
    nrm_all = 0.0
    for corner in corners:
        o = [corner[d] * n[d] for d in range(3)]          # block origin (x, y, z)
        mode = rank_mode(corner)
        lb = [cut(x, n, o) for x in b]; lrhs = cut(rhs, n, o); ldinv = cut(dinvN, n, o)
        lphi = cut(phi, n, o); lc = cut(cphi, cn, [x // 2 for x in o])
        out, crhs, czero = view(corner)
        delta = None
        if p2p:
            delta = (C.c_long * 27)()
            for q in range(27):
                off = (q % 3 - 1, (q // 3) % 3 - 1, q // 9 - 1)
                pc = [corner[d] + off[d] for d in range(3)]
                ok = True
                for d in range(3):
                    if off[d] == 0:
                        continue
                    if d not in split:
                        ok = False
                    elif pc[d] < 0 or pc[d] >= pgrid[d]:
                        if per[d]:
                            pc[d] %= pgrid[d]
                        else:
                            ok = False
                if ok:
                    delta[q] = heap[tuple(pc)].ctypes.data - heap[corner].ctypes.data
        nrm = np.zeros(1)
        fn = emu.emu_sweep3_p2p
        rc = fn(1, pre, post, cfg, (C.c_int * 3)(*n), (C.c_int * 6)(*[m for d in mode for m in d]), sum(o) & 1, Pp(h2),
                Pp(lrhs), Pp(lb[0]), Pp(lb[1]), Pp(lb[2]), Pp(lphi), Pp(out), Pp(lc), Pp(crhs), Pp(czero), Pp(nrm), 8, PAD, delta, Pp(ldinv))
        assert rc == 0
        nrm_all = max(nrm_all, nrm[0])
    for corner in corners:
        o = [corner[d] * n[d] for d in range(3)]
        mode = rank_mode(corner)
        out, crhs, czero = view(corner)
        # own cells, and in peer-memory mode the 3 ghost layers on the sides shared with another rank
        g = [[(3 if (p2p and mode[d][s] == M_GHOST) else 0) for s in range(2)] for d in range(3)]
        R = tuple(slice(PAD - g[d][0], PAD + n[d] + g[d][1]) for d in (2, 1, 0))
        RN = tuple(slice(o[d] + PAD - g[d][0], o[d] + PAD + n[d] + g[d][1]) for d in (2, 1, 0))
        want = refg[RN]
        assert np.abs(out[R] - want).max() <= 1e-12 * np.abs(want).max(), corner
        if post == 2:
            co = [x // 2 for x in o]
            CR = tuple(slice(PAD - g[d][0], PAD + cn[d] + g[d][1]) for d in (2, 1, 0))
            CRN = tuple(slice(co[d] + PAD - g[d][0], co[d] + PAD + cn[d] + g[d][1]) for d in (2, 1, 0))
            assert np.abs(crhs[CR] - cres[CRN]).max() <= 1e-10 * np.abs(res).max(), corner
            assert np.all(czero[CR] == 0.0), corner
    if post == 3:
        assert abs(nrm_all - np.abs(res).max()) <= 1e-10 * np.abs(res).max()
