"""
CPU execution of the REAL Godunov kernel source + stage orchestration (varden_b200/csrc/vdn_godunov_kernels.cuh) under
tests/emu/cuda_emu.h, checked BIT FOR BIT against the CPU oracle (oracle/, pinned to the reference routines
velpred_3d velpred.f90:1776 and mkflux_3d mkflux.f90:1186).  The launch structure is one direction of one stage per launch
(staged): the 2-D product path, run here in its 3-D instantiation as an independent second implementation of what the
plane-marching kernels (tests/test_emu_march.py) compute.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")
W, NS, IN, OUT, PER, SYM = O.SLIP_WALL, O.NO_SLIP_WALL, O.INLET, O.OUTLET, O.PERIODIC, O.SYMMETRY
FOEXTRAP, EXT_DIR, HOEXTRAP, REFLECT_ODD, REFLECT_EVEN, INTERIOR = 22, 23, 24, 20, 21, 0


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(EMU, "libemu_godunov.so")
    csrc = os.path.join(HERE, "..", "varden_b200", "csrc")
    src = [os.path.join(EMU, "emu_godunov.cpp"), os.path.join(EMU, "cuda_emu.h"), os.path.join(csrc, "vdn_godunov_march.cuh"),
           os.path.join(csrc, "vdn_godunov_kernels.cuh"), os.path.join(csrc, "vdn_common.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["g++", "-O1", "-std=c++20", "-pthread", "-fPIC", "-shared", "-ffp-contract=off", src[0], "-o", so])
    return C.CDLL(so)


def adv_bc_table(phys_bc, dim, nscal):
    """define_bc_tower.f90:158-340 for the velocity and scalar comps"""
    t = np.zeros((dim + nscal, 3, 2), dtype=np.int32)
    for d in range(dim):
        for s in range(2):
            p = int(phys_bc[d][s])
            if p == W:
                t[:dim, d, s] = HOEXTRAP; t[d, d, s] = EXT_DIR; t[dim:, d, s] = HOEXTRAP
            elif p == NS:
                t[:dim, d, s] = EXT_DIR; t[dim:, d, s] = HOEXTRAP
            elif p == IN:
                t[:, d, s] = EXT_DIR
            elif p == OUT:
                t[:, d, s] = FOEXTRAP
            elif p == SYM:
                t[:dim, d, s] = REFLECT_EVEN; t[d, d, s] = REFLECT_ODD; t[dim:, d, s] = REFLECT_EVEN
    return t


CASES = {
    "rt": lambda: O.rt_state(16, dim=3, max_grid_size=16),
    "mixed": lambda: O.random_state([16, 12, 20], dim=3, max_grid_size=32, phys_bc=[[IN, OUT], [NS, W], [PER, PER]], seed=2),
    "outx_so2": lambda: O.random_state([12, 16, 12], dim=3, max_grid_size=32, phys_bc=[[OUT, IN], [PER, PER], [W, OUT]], seed=3,
                                       params=O.Params(dim=3, slope_order=2)),
    "minion": lambda: O.random_state(12, dim=3, max_grid_size=32, phys_bc=[[W, W], [SYM, NS], [IN, OUT]], seed=5,
                                     params=O.Params(dim=3, use_minion=True)),
}


def P_(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("case", sorted(CASES))
def test_godunov_kernels_bit_exact(emu, case, fused=0):
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

    # ---- velpred ----
    u = st["uold"][0]; force = ref["vel_force_1"][0]
    umax = np.abs(O.valid(geom, u, 0, 3)).max()
    eps = 1e-8 if umax == 0 else 1e-8 * umax
    um = [np.full_like(ref["umac_pred"][d][0], 1.0e20) for d in range(3)]
    vt = np.ascontiguousarray(tab[:3]).astype(np.int32)
    rc = emu.emu_velpred(fused, nA, pbc, vt.ctypes.data_as(C.c_void_p), P.slope_order, P.use_minion, dbl(dt), h, dbl(eps),
                         P_(u), P_(force), P_(um[0]), P_(um[1]), P_(um[2]))
    assert rc == 0
    for d in range(3):
        a, b = O.valid(geom, um[d], 0, 1, d), O.valid(geom, ref["umac_pred"][d][0], 0, 1, d)
        assert np.array_equal(a, b), ("umac", d, np.abs(a - b).max())

    # ---- mkflux: scalars, then velocity, with the oracle's projected MAC velocities ----
    mac = [np.asfortranarray(ref["umac"][d][0][..., 0]) for d in range(3)]
    fmax = max(np.abs(O.valid(geom, ref["umac"][d][0], 0, 1, d)).max() for d in range(3))
    eps = 1e-8 if fmax == 0 else 1e-8 * fmax
    zero_rhs = np.zeros(geom.box_shape(0, 1), order='F')
    for is_vel in (0, 1):
        src = st["uold"][0] if is_vel else st["sold"][0]
        frc = ref["vel_force_1"][0] if is_vel else ref["scal_force_1"][0]
        want = ref["uedge"] if is_vel else ref["sedge"]
        for comp in range(dim if is_vel else nscal):
            cons = int(not is_vel and comp == 0)
            s = np.asfortranarray(src[..., comp]); f = np.asfortranarray(frc[..., comp])
            se = [np.full(want[d][0].shape[:3], np.nan, order='F') for d in range(3)]
            fl = [np.full(want[d][0].shape[:3], np.nan, order='F') for d in range(3)]
            sb = np.ascontiguousarray(tab[(0 if is_vel else dim) + comp]).astype(np.int32)
            rc = emu.emu_mkflux(fused, nA, pbc, sb.ctypes.data_as(C.c_void_p), P.slope_order, P.use_minion, is_vel, comp, cons, is_vel,
                                dbl(dt), h, dbl(eps), P_(s), P_(mac[0]), P_(mac[1]), P_(mac[2]), P_(f), P_(zero_rhs),
                                P_(se[0]), P_(se[1]), P_(se[2]), P_(fl[0]), P_(fl[1]), P_(fl[2]))
            assert rc == 0
            for d in range(3):
                assert np.array_equal(se[d], want[d][0][..., comp]), ("edge", is_vel, comp, d, np.nanmax(np.abs(se[d] - want[d][0][..., comp])))
                if cons:
                    assert np.array_equal(fl[d], ref["sflux"][d][0][..., comp]), ("flux", comp, d)
