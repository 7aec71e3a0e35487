"""
CPU execution of the REAL plane-marching Godunov kernels (varden_b200/csrc/vdn_godunov_march.cuh: k_mkflux_march,
k_velpred_march -- warp shuffles, shared-memory exchange, z-chunking and all) under tests/emu/cuda_emu.h, checked BIT FOR
BIT against the CPU oracle (oracle/, pinned to velpred_3d velpred.f90:1776 and mkflux_3d mkflux.f90:1186).

The emulation builds use a 32 x 6 thread tile (30 x 4 output columns) so that small grids still span several tiles in x and y,
and a small resident-CTA count so that the z range is cut into several chunks (each with its own warm-up planes).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from test_emu_godunov import adv_bc_table, P_, W, NS, IN, OUT, PER, SYM

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")


def build(ncg):
    dev = os.environ.get("VDN_MARCH_HEADER")                     # development: test a working copy of the kernel header
    so = os.path.join(EMU, "libemu_march_ncg%d%s.so" % (ncg, "_dev" if dev else ""))
    csrc = os.path.join(HERE, "..", "varden_b200", "csrc")
    src = [os.path.join(EMU, "emu_godunov.cpp"), os.path.join(EMU, "cuda_emu.h"), os.path.join(csrc, "vdn_godunov_march.cuh"),
           os.path.join(csrc, "vdn_godunov_kernels.cuh"), os.path.join(csrc, "vdn_common.cuh")]
    if dev:
        src.append(dev)
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["g++", "-O1", "-std=c++20", "-pthread", "-fPIC", "-shared", "-ffp-contract=off",
                               "-DMARCH_TYT=6", "-DMARCH_NCG=%d" % ncg] + (['-DMARCH_HEADER="%s"' % dev] if dev else []) + [src[0], "-o", so])
    return C.CDLL(so)


@pytest.fixture(scope="module", params=[1, 3])
def emu(request):
    return build(request.param)


CASES = {
    # periodic x,y / walls z (the bench problem), 2 x 4 tiles, 1 chunk
    "rt": lambda: O.rt_state([32, 12, 8], dim=3, max_grid_size=64),
    # every spacing a power of two: the instantiation with x / h compiled as the exact product x * (1/h)
    "rt_pow2": lambda: O.rt_state([32, 8, 16], dim=3, max_grid_size=64),
    "mixed_pow2": lambda: O.random_state([16, 8, 8], dim=3, max_grid_size=64, phys_bc=[[W, IN], [PER, PER], [OUT, NS]], seed=21,
                                         prob_hi=[1.0, 0.5, 0.5]),
    # every override type, non-cubic, tiles straddle every boundary
    "mixed": lambda: O.random_state([34, 10, 12], dim=3, max_grid_size=64, phys_bc=[[IN, OUT], [NS, W], [PER, PER]], seed=2),
    "outx_so2": lambda: O.random_state([12, 16, 12], dim=3, max_grid_size=32, phys_bc=[[OUT, IN], [PER, PER], [W, OUT]], seed=3,
                                       params=O.Params(dim=3, slope_order=2)),
    "minion": lambda: O.random_state([12, 8, 10], dim=3, max_grid_size=32, phys_bc=[[W, W], [SYM, NS], [IN, OUT]], seed=5,
                                     params=O.Params(dim=3, use_minion=True)),
    "so0": lambda: O.random_state([8, 8, 8], dim=3, max_grid_size=32, phys_bc=[[PER, PER], [W, W], [NS, NS]], seed=7,
                                  params=O.Params(dim=3, slope_order=0)),
    # n0 = 32: the last lane of the first tile row sits at cell n0-2, whose upper fromm is the one-sided hi slope
    "zwalls_n32": lambda: O.random_state([32, 9, 20], dim=3, max_grid_size=64, phys_bc=[[NS, IN], [W, OUT], [OUT, NS]], seed=11),
}


@pytest.mark.parametrize("slots", [1, 1000])
@pytest.mark.parametrize("case", sorted(CASES))
def test_march_kernels_bit_exact(emu, case, slots):
    geom, P, st, dt = CASES[case]()
    assert geom.nboxes == 1
    dim, nscal = 3, P.nscal
    n = [geom.dhi[d] - geom.dlo[d] + 1 for d in range(3)]
    ref = O.stagewise(geom, P, st, dt, mac_rel_eps=1e-11)
    nA = (C.c_int * 3)(*n)
    pbc = (C.c_int * 6)(*[int(x) for x in np.asarray(geom.phys_bc).ravel()[:6]])
    h = (C.c_double * 3)(*geom.dx[:3])
    tab = adv_bc_table(geom.phys_bc, dim, nscal)
    dbl = C.c_double

    if hasattr(emu, "emu_velpred_march"):
        u = st["uold"][0]; force = ref["vel_force_1"][0]
        umax = np.abs(O.valid(geom, u, 0, 3)).max()
        eps = 1e-8 if umax == 0 else 1e-8 * umax
        um = [np.full_like(ref["umac_pred"][d][0], 1.0e20) for d in range(3)]
        vt = np.ascontiguousarray(tab[:3]).astype(np.int32)
        rc = emu.emu_velpred_march(nA, pbc, vt.ctypes.data_as(C.c_void_p), P.slope_order, P.use_minion, dbl(dt), h, dbl(eps), slots,
                                   P_(u), P_(force), P_(um[0]), P_(um[1]), P_(um[2]))
        assert rc == 0
        for d in range(3):
            a, b = O.valid(geom, um[d], 0, 1, d), O.valid(geom, ref["umac_pred"][d][0], 0, 1, d)
            assert np.array_equal(a, b), ("umac", d, np.abs(a - b).max(), np.argwhere(a != b)[:5])

    mac = [np.asfortranarray(ref["umac"][d][0][..., 0]) for d in range(3)]
    fmax = max(np.abs(O.valid(geom, ref["umac"][d][0], 0, 1, d)).max() for d in range(3))
    eps = 1e-8 if fmax == 0 else 1e-8 * fmax
    for is_vel in (0, 1):
        src = np.asfortranarray(st["uold"][0] if is_vel else st["sold"][0])
        frc = np.asfortranarray(ref["vel_force_1"][0] if is_vel else ref["scal_force_1"][0])
        want = ref["uedge"] if is_vel else ref["sedge"]
        ncomp = dim if is_vel else nscal
        se = [np.full(want[d][0].shape, np.nan, order='F') for d in range(3)]
        fl = [np.full(want[d][0].shape[:3], np.nan, order='F') for d in range(3)]
        sb = np.ascontiguousarray(tab[(0 if is_vel else dim):(0 if is_vel else dim) + ncomp]).astype(np.int32)
        rc = emu.emu_mkflux_march(nA, pbc, sb.ctypes.data_as(C.c_void_p), P.slope_order, P.use_minion, is_vel, ncomp, dbl(dt), h, dbl(eps), slots,
                                  P_(src), P_(mac[0]), P_(mac[1]), P_(mac[2]), P_(frc),
                                  P_(se[0]), P_(se[1]), P_(se[2]), P_(fl[0]), P_(fl[1]), P_(fl[2]))
        assert rc == 0
        for d in range(3):
            for comp in range(ncomp):
                a, b = se[d][..., comp], want[d][0][..., comp]
                assert np.array_equal(a, b), ("edge", is_vel, comp, d, np.nanmax(np.abs(a - b)), np.argwhere(a != b)[:5])
            if not is_vel:
                a, b = fl[d], ref["sflux"][d][0][..., 0]
                assert np.array_equal(a, b), ("flux", d, np.argwhere(a != b)[:5])
