"""
oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front end of the CPU oracle (oracle/liboracle.so, built by oracle/Makefile)
plus the synthetic problem set-up shared by tests/ and bench.py's cpu_baseline leg.
Nothing under varden_b200/ imports this module.

Array convention: every box is a numpy array of shape (n0, n1, n2, ncomp), order='F',
covering lo-ng .. hi+ng (+1 in the face direction); 2-D uses n2 == 1.
"""
import ctypes as C
import os
import subprocess
import sys
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
if os.path.dirname(_HERE) not in sys.path:
    sys.path.insert(0, os.path.dirname(_HERE))

# FBoxLib bc_module codes (see orc_common.h)
PERIODIC, INTERIOR, INLET, OUTLET, SYMMETRY, SLIP_WALL, NO_SLIP_WALL = -1, 0, 11, 12, 13, 14, 15
REFLECT_ODD, REFLECT_EVEN, FOEXTRAP, EXT_DIR, HOEXTRAP = 20, 21, 22, 23, 24


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


_lib = None
_timing = False


def use_timing_build(threads=None):
    """bench.py only: load the -O3 -march=native build (compiled HERE, on the machine that runs it) instead of the parity build, and
    set the OpenMP thread count explicitly (torchrun exports OMP_NUM_THREADS=1).  Must be called before the first lib() use."""
    global _lib, _timing
    assert _lib is None, "use_timing_build must come before the first oracle call"
    if threads:
        os.environ["OMP_NUM_THREADS"] = str(threads)
    # always rebuilt: -march=native must match the machine that runs it, and a snapshot may carry a build from another host
    subprocess.check_call(["make", "-B", "-C", _HERE, "fast"], stdout=subprocess.DEVNULL)
    _lib = C.CDLL(os.path.join(_HERE, "_fast", "liboracle_fast.so"))
    _timing = True
    if threads:
        _lib.orc_set_threads(int(threads))
    return int(_lib.orc_get_threads())



def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


class OrcParams(C.Structure):
    _fields_ = [("dim", C.c_int), ("nscal", C.c_int), ("slope_order", C.c_int), ("use_minion", C.c_int),
                ("boussinesq", C.c_int), ("stencil_order", C.c_int),
                ("visc_coef", C.c_double), ("diff_coef", C.c_double),
                ("bcval", C.c_double * 30),
                ("mg_rel_eps", C.c_double), ("mg_bottom_eps", C.c_double),
                ("mg_max_cycles", C.c_int), ("mg_nu1", C.c_int), ("mg_nu2", C.c_int), ("mg_verbose", C.c_int)]


class OrcGeom(C.Structure):
    _fields_ = [("nboxes", C.c_int), ("blo", C.POINTER(C.c_int)), ("bhi", C.POINTER(C.c_int)),
                ("dlo", C.c_int * 3), ("dhi", C.c_int * 3), ("phys_bc", C.c_int * 6), ("dx", C.c_double * 3)]


class Params:
    """The probin values the path reads (src/_parameters)."""

    def __init__(self, dim=3, nscal=2, slope_order=4, use_minion=False, boussinesq=0, stencil_order=2,
                 visc_coef=0.0, diff_coef=0.0, bcval=None, mg_rel_eps=1e-10, mg_bottom_eps=1e-3,
                 mg_max_cycles=100, mg_nu1=2, mg_nu2=2, mg_verbose=0):
        self.dim, self.nscal, self.slope_order, self.use_minion = dim, nscal, slope_order, int(use_minion)
        self.boussinesq, self.stencil_order = boussinesq, stencil_order
        self.visc_coef, self.diff_coef = visc_coef, diff_coef
        self.bcval = np.zeros((5, 3, 2)) if bcval is None else np.asarray(bcval, dtype=float).reshape(5, 3, 2)
        self.mg_rel_eps, self.mg_bottom_eps = mg_rel_eps, mg_bottom_eps
        self.mg_max_cycles, self.mg_nu1, self.mg_nu2, self.mg_verbose = mg_max_cycles, mg_nu1, mg_nu2, mg_verbose

    def to_c(self):
        p = OrcParams()
        p.dim, p.nscal, p.slope_order, p.use_minion = self.dim, self.nscal, self.slope_order, self.use_minion
        p.boussinesq, p.stencil_order = self.boussinesq, self.stencil_order
        p.visc_coef, p.diff_coef = self.visc_coef, self.diff_coef
        for i, v in enumerate(self.bcval.ravel()):
            p.bcval[i] = v
        p.mg_rel_eps, p.mg_bottom_eps = self.mg_rel_eps, self.mg_bottom_eps
        p.mg_max_cycles, p.mg_nu1, p.mg_nu2, p.mg_verbose = self.mg_max_cycles, self.mg_nu1, self.mg_nu2, self.mg_verbose
        return p


from varden_b200.problems import Geom, chop_domain, mf_alloc, valid, rt_problem  # noqa: E402  (shared problem set-up)


def _geom_to_c(self):
    g = OrcGeom()
    g.nboxes = self.nboxes
    g.blo = self._blo.ctypes.data_as(C.POINTER(C.c_int))
    g.bhi = self._bhi.ctypes.data_as(C.POINTER(C.c_int))
    for d in range(3):
        g.dlo[d], g.dhi[d], g.dx[d] = self.dlo[d], self.dhi[d], self.dx[d]
    for i, v in enumerate(self.phys_bc.ravel()):
        g.phys_bc[i] = int(v)
    return g


Geom.to_c = _geom_to_c


def _pp(mf):
    """list of numpy arrays -> double**"""
    if mf is None:
        return None
    arr = (C.c_void_p * len(mf))(*[a.ctypes.data for a in mf])
    return C.cast(arr, C.POINTER(C.POINTER(C.c_double)))


def valid(geom, a, ib, ng, face_dir=-1):
    """view of the valid region of box ib"""
    sl = []
    for d in range(3):
        if d < geom.dim:
            n = a.shape[d]
            sl.append(slice(ng, n - ng))
        else:
            sl.append(slice(None))
    return a[tuple(sl)]


def fill_boundary(geom, mf, ng, ncomp, face_dir=-1):
    g = geom.to_c()
    lib().orc_fill_boundary(C.byref(g), C.c_int(geom.dim), _pp(mf), C.c_int(ng), C.c_int(ncomp), C.c_int(face_dir))


def fill_and_physbc(geom, params, mf, ng, ncomp_total, scomp, bccomp, nc, same_boundary=False):
    g, p = geom.to_c(), params.to_c()
    lib().orc_fill_and_physbc(C.byref(g), C.byref(p), _pp(mf), C.c_int(ng), C.c_int(ncomp_total), C.c_int(scomp),
                              C.c_int(bccomp), C.c_int(nc), C.c_int(int(same_boundary)))


def advance(geom, params, st, dt, mac_rel_eps=-1.0, want_phi=True):
    """One pass of the hot path (advance_timestep.f90:66-124).  st: dict with uold,sold,gp,ext_vel_force,ext_scal_force[,lapu]."""
    dim, nscal = geom.dim, params.nscal
    g, p = geom.to_c(), params.to_c()
    lapu = st.get("lapu") or mf_alloc(geom, 0, dim)
    out = dict(unew=mf_alloc(geom, 3, dim), snew=mf_alloc(geom, 3, nscal), rhohalf=mf_alloc(geom, 1, 1),
               umac=[mf_alloc(geom, 1, 1, d) for d in range(dim)], phi=mf_alloc(geom, 1, 1) if want_phi else None)
    um = out["umac"] + [None] * (3 - dim)
    res = C.c_double(0.0)
    f = lib().orc_advance_mf
    f.restype = C.c_int
    cycles = f(C.byref(g), C.byref(p), _pp(st["uold"]), _pp(st["sold"]), _pp(st["gp"]), _pp(st["ext_vel_force"]),
               _pp(st["ext_scal_force"]), _pp(lapu), _pp(out["unew"]), _pp(out["snew"]), _pp(out["rhohalf"]),
               _pp(um[0]), _pp(um[1]), _pp(um[2]), _pp(out["phi"]), C.c_double(dt), C.c_double(mac_rel_eps), C.byref(res))
    out["mac_cycles"], out["mac_resnorm"] = cycles, res.value
    return out


# ---------------------------------------------------------------------------------------------
# synthetic problems; initial data follows src/initdata.f90:195-200,261-274
# ---------------------------------------------------------------------------------------------
def rt_state(n, dim=3, max_grid_size=256, ratio=2.0, grav=-9.8, seeded_velocity=True, params=None, phys_bc=None):
    """rt_problem (varden_b200/problems.py) + the path-boundary ghost fill of varden.f90:291-300."""
    if params is None:
        params = Params(dim=dim, nscal=2)
    geom, st, dt = rt_problem(n, dim=dim, max_grid_size=max_grid_size, ratio=ratio, grav=grav, nscal=params.nscal,
                              seeded_velocity=seeded_velocity, phys_bc=phys_bc)
    fill_and_physbc(geom, params, st["uold"], 3, dim, 0, 0, dim)
    fill_and_physbc(geom, params, st["sold"], 3, params.nscal, 0, dim, params.nscal)
    fill_boundary(geom, st["gp"], 1, dim)
    return geom, params, st, dt


def bubble_state(n=64, max_grid_size=32, params=None):
    """bubble_problem (BASELINE config 1: exec/test/inputs_2d-regt, max_levs = 1) + the path-boundary ghost fill of varden.f90:291-300."""
    from varden_b200.problems import bubble_problem
    if params is None:
        params = Params(dim=2, nscal=2)
    geom, st, dt = bubble_problem(n, max_grid_size=max_grid_size, nscal=params.nscal)
    fill_and_physbc(geom, params, st["uold"], 3, 2, 0, 0, 2)
    fill_and_physbc(geom, params, st["sold"], 3, params.nscal, 0, 2, params.nscal)
    fill_boundary(geom, st["gp"], 1, 2)
    return geom, params, st, dt


def random_state(n, dim=3, max_grid_size=256, phys_bc=None, seed=0, params=None, umag=1.0, prob_hi=None):
    """Randomised but smooth-ish state exercising all selects; any phys_bc combination."""
    rng = np.random.default_rng(seed)
    if np.isscalar(n):
        n = [n] * dim
    if phys_bc is None:
        phys_bc = [[SLIP_WALL, SLIP_WALL]] * dim
    geom = Geom(dim, n, phys_bc, max_grid_size=max_grid_size) if prob_hi is None else Geom(dim, n, phys_bc, prob_hi=prob_hi, max_grid_size=max_grid_size)
    if params is None:
        bcval = np.zeros((5, 3, 2))
        bcval[0:3] = rng.uniform(-0.5, 0.5, size=(3, 3, 2)) * umag     # inflow velocities
        bcval[3] = rng.uniform(1.0, 2.0, size=(3, 2))                  # rho_bc
        bcval[4] = rng.uniform(0.0, 1.0, size=(3, 2))                  # trac_bc
        params = Params(dim=dim, nscal=2, bcval=bcval)
    nn = [geom.n_cell[d] for d in range(3)]
    # global random fields (smoothed a little so that slopes/limiters take both branches)
    def field(scale, offset=0.0):
        f = rng.standard_normal(nn)
        for d in range(dim):
            f = 0.5 * f + 0.25 * (np.roll(f, 1, d) + np.roll(f, -1, d))
        return offset + scale * f
    U = [field(umag) for _ in range(dim)]
    RHO = np.abs(field(0.5, 1.5)) + 0.2
    TR = field(1.0)
    GP = [field(0.3) for _ in range(dim)]
    st = dict(uold=mf_alloc(geom, 3, dim), sold=mf_alloc(geom, 3, params.nscal), gp=mf_alloc(geom, 1, dim),
              ext_vel_force=mf_alloc(geom, 1, dim), ext_scal_force=mf_alloc(geom, 1, params.nscal))
    for ib, (lo, hi) in enumerate(geom.boxes):
        sl = tuple(slice(lo[d], hi[d] + 1) for d in range(3))
        for c in range(dim):
            valid(geom, st["uold"][ib], ib, 3)[..., c] = U[c][sl]
            valid(geom, st["gp"][ib], ib, 1)[..., c] = GP[c][sl]
        valid(geom, st["sold"][ib], ib, 3)[..., 0] = RHO[sl]
        valid(geom, st["sold"][ib], ib, 3)[..., 1] = TR[sl]
        st["ext_vel_force"][ib][..., dim - 1] = -9.8
        st["ext_scal_force"][ib][..., 1] = 0.1
    fill_and_physbc(geom, params, st["uold"], 3, dim, 0, 0, dim)
    fill_and_physbc(geom, params, st["sold"], 3, params.nscal, 0, dim, params.nscal)
    fill_boundary(geom, st["gp"], 1, dim)
    # gp: the reference only fill_boundary's it (varden.f90:294); give the physical ghosts finite values
    fill_and_physbc(geom, params, st["gp"], 1, dim, 0, dim + params.nscal + 1, dim, same_boundary=True)
    umax = max(np.abs(u).max() for u in U)
    dt = 0.4 * min(geom.dx[:dim]) / max(umax, 1e-3)
    return geom, params, st, dt


# ---------------------------------------------------------------------------------------------
# stage-wise entry points (same call sequence as the reference module procedures)
# ---------------------------------------------------------------------------------------------
def _um3(umac, dim):
    arr = (C.POINTER(C.POINTER(C.c_double)) * 3)()
    for d in range(3):
        arr[d] = _pp(umac[d]) if d < dim else None
    return arr


def mkvelforce(geom, params, vel_force, ext, gp, s, ng_s, ncomp_s, lapu, visc_fac):
    g, p = geom.to_c(), params.to_c()
    lib().orc_mkvelforce_mf(C.byref(g), C.byref(p), _pp(vel_force), _pp(ext), _pp(gp), _pp(s), C.c_int(ng_s), C.c_int(ncomp_s),
                            _pp(lapu), C.c_double(visc_fac))


def mkscalforce(geom, params, scal_force, ext, laps, diff_fac):
    g, p = geom.to_c(), params.to_c()
    lib().orc_mkscalforce_mf(C.byref(g), C.byref(p), _pp(scal_force), _pp(ext), _pp(laps), C.c_double(diff_fac))


def velpred(geom, params, u, umac, force, dt):
    g, p = geom.to_c(), params.to_c()
    lib().orc_velpred_mf(C.byref(g), C.byref(p), _pp(u), _um3(umac, geom.dim), _pp(force), C.c_double(dt))


def macproject(geom, params, umac, rho, mac_rhs, rel_eps=-1.0, want_phi=True):
    g, p = geom.to_c(), params.to_c()
    phi = mf_alloc(geom, 1, 1) if want_phi else None
    res = C.c_double(0.0)
    f = lib().orc_macproject_mf
    f.restype = C.c_int
    cyc = f(C.byref(g), C.byref(p), _um3(umac, geom.dim), _pp(rho), C.c_int(params.nscal), _pp(mac_rhs), _pp(phi),
            C.c_double(rel_eps), C.byref(res))
    return cyc, res.value, phi


def mkflux(geom, params, sold, ncomp, sedge, flux, umac, force, mac_rhs, dt, is_vel, is_cons):
    g, p = geom.to_c(), params.to_c()
    ic = (C.c_int * 8)(*([int(x) for x in is_cons] + [0] * (8 - len(is_cons))))
    lib().orc_mkflux_mf(C.byref(g), C.byref(p), _pp(sold), C.c_int(ncomp), _um3(sedge, geom.dim), _um3(flux, geom.dim),
                        _um3(umac, geom.dim), _pp(force), _pp(mac_rhs), C.c_double(dt), C.c_int(int(is_vel)), ic)


def update(geom, params, sold, ncomp, umac, sedge, flux, force, snew, dt, is_vel, is_cons):
    g, p = geom.to_c(), params.to_c()
    ic = (C.c_int * 8)(*([int(x) for x in is_cons] + [0] * (8 - len(is_cons))))
    lib().orc_update_mf(C.byref(g), C.byref(p), _pp(sold), C.c_int(ncomp), _um3(umac, geom.dim), _um3(sedge, geom.dim),
                        _um3(flux, geom.dim), _pp(force), _pp(snew), C.c_double(dt), C.c_int(int(is_vel)), ic)


def make_at_halftime(geom, params, rhohalf, sold, snew):
    for ib, (lo, hi) in enumerate(geom.boxes):
        lo_a = (C.c_int * 3)(*lo)
        hi_a = (C.c_int * 3)(*hi)
        lib().orc_make_at_halftime(rhohalf[ib].ctypes.data_as(C.POINTER(C.c_double)), sold[ib].ctypes.data_as(C.POINTER(C.c_double)),
                                   snew[ib].ctypes.data_as(C.POINTER(C.c_double)), lo_a, hi_a, C.c_int(geom.dim), C.c_int(1), C.c_int(3))
    fill_and_physbc(geom, params, rhohalf, 1, 1, 0, geom.dim, 1)


def stagewise(geom, params, st, dt, mac_rel_eps=-1.0):
    """The whole path with every intermediate kept (what the stage-wise GPU parity tests compare against)."""
    dim, nscal = geom.dim, params.nscal
    lapu = st.get("lapu") or mf_alloc(geom, 0, dim)
    o = {}
    o["vel_force_1"] = mf_alloc(geom, 1, dim)
    mkvelforce(geom, params, o["vel_force_1"], st["ext_vel_force"], st["gp"], st["sold"], 3, nscal, lapu, 1.0)
    o["umac_pred"] = [mf_alloc(geom, 1, 1, d, val=1.0e20) for d in range(dim)]
    velpred(geom, params, st["uold"], o["umac_pred"], o["vel_force_1"], dt)
    o["umac"] = [[a.copy(order='F') for a in o["umac_pred"][d]] for d in range(dim)]
    mac_rhs = mf_alloc(geom, 1, 1)
    o["mac_cycles"], o["mac_resnorm"], o["phi"] = macproject(geom, params, o["umac"], st["sold"], mac_rhs, mac_rel_eps)
    o["scal_force_1"] = mf_alloc(geom, 1, nscal)
    laps = mf_alloc(geom, 0, nscal)
    mkscalforce(geom, params, o["scal_force_1"], st["ext_scal_force"], laps, 1.0)
    o["sedge"] = [mf_alloc(geom, 0, nscal, d) for d in range(dim)]
    o["sflux"] = [mf_alloc(geom, 0, nscal, d) for d in range(dim)]
    divu = mf_alloc(geom, 1, 1)
    is_cons_s = [1] + [0] * (nscal - 1)
    mkflux(geom, params, st["sold"], nscal, o["sedge"], o["sflux"], o["umac"], o["scal_force_1"], divu, dt, False, is_cons_s)
    o["scal_force_2"] = mf_alloc(geom, 1, nscal)
    mkscalforce(geom, params, o["scal_force_2"], st["ext_scal_force"], laps, 0.0)
    o["snew"] = mf_alloc(geom, 3, nscal)
    update(geom, params, st["sold"], nscal, o["umac"], o["sedge"], o["sflux"], o["scal_force_2"], o["snew"], dt, False, is_cons_s)
    o["rhohalf"] = mf_alloc(geom, 1, 1)
    make_at_halftime(geom, params, o["rhohalf"], st["sold"], o["snew"])
    o["uedge"] = [mf_alloc(geom, 0, dim, d) for d in range(dim)]
    uflux = [mf_alloc(geom, 0, dim, d) for d in range(dim)]
    mkflux(geom, params, st["uold"], dim, o["uedge"], uflux, o["umac"], o["vel_force_1"], mac_rhs, dt, True, [0] * dim)
    o["vel_force_2"] = mf_alloc(geom, 1, dim)
    mkvelforce(geom, params, o["vel_force_2"], st["ext_vel_force"], st["gp"], o["rhohalf"], 1, 1, lapu, 0.0)
    o["unew"] = mf_alloc(geom, 3, dim)
    update(geom, params, st["uold"], dim, o["umac"], o["uedge"], uflux, o["vel_force_2"], o["unew"], dt, True, [0] * dim)
    return o


# ---------------------------------------------------------------------------------------------
# macproject glue, stage by stage (macproject.f90:137-225, :280-336, :403-505) -- used by the reference-pin tests
# ---------------------------------------------------------------------------------------------
def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ia(v):
    return (C.c_int * 3)(*[int(x) for x in v])


def _box_bc(geom, ib):
    g = geom.to_c()
    pb, ell = (C.c_int * 6)(), (C.c_int * 6)()
    lib().orc_box_phys_bc(C.byref(g), C.c_int(ib), pb)
    lib().orc_ell_bc_press(pb, C.c_int(geom.dim), ell)
    return pb, ell


def divumac(geom, umac, mac_rhs, rh):
    """rh = mac_rhs - div(umac)"""
    dim = geom.dim
    dx = (C.c_double * 3)(*geom.dx)
    for ib, (lo, hi) in enumerate(geom.boxes):
        lib().orc_divumac(_dp(umac[0][ib]), _dp(umac[1][ib]), _dp(umac[2][ib]) if dim == 3 else None, C.c_int(1),
                          _dp(mac_rhs[ib]), C.c_int(1), _dp(rh[ib]), C.c_int(0), dx, _ia(lo), _ia(hi), C.c_int(dim))


def mk_mac_coeffs(geom, rho, ng_r, beta):
    dim = geom.dim
    for ib, (lo, hi) in enumerate(geom.boxes):
        lib().orc_mk_mac_coeffs(_dp(beta[0][ib]), _dp(beta[1][ib]), _dp(beta[2][ib]) if dim == 3 else None, C.c_int(0),
                                _dp(rho[ib]), C.c_int(ng_r), _ia(lo), _ia(hi), C.c_int(dim))


def estdt(geom, u, ng_u, s, ng_s, gp, ng_g, ext_vel_force, ng_f, dtold=-1.0, cflfac=0.5, max_dt_growth=1.1):
    """estdt.f90:15-87: dt for the next step from the new velocity, density, grad(p) and the external force (probin defaults: cflfac 0.5,
    max_dt_growth 1.1, src/_parameters)"""
    dim = geom.dim
    nb = geom.nboxes
    blo = (C.c_int * (3 * nb))(*[int(x) for lo, hi in geom.boxes for x in list(lo) + [0] * (3 - len(lo))])
    bhi = (C.c_int * (3 * nb))(*[int(x) for lo, hi in geom.boxes for x in list(hi) + [0] * (3 - len(hi))])
    dx = (C.c_double * 3)(*(list(geom.dx) + [1.0] * (3 - len(geom.dx))))
    f = lib().orc_estdt_mf
    f.restype = C.c_double
    return f(_pp(u), C.c_int(ng_u), _pp(s), C.c_int(ng_s), _pp(gp), C.c_int(ng_g), _pp(ext_vel_force), C.c_int(ng_f), blo, bhi, C.c_int(nb),
             C.c_int(dim), dx, C.c_double(dtold), C.c_double(cflfac), C.c_double(max_dt_growth))


# ---------------------------------------------------------------------------------------------
# visc_solve / diff_scalar_solve (viscsolve.f90) -- ** parity unpinned ** like the MAC solve (F_MG absent)
# ---------------------------------------------------------------------------------------------
def _gather(geom, mf, ng, comp, grow=0):
    """whole-domain array of one component: valid cells of every box, plus `grow` layers around the domain taken from the boxes' ghost cells"""
    dim = geom.dim
    N = [geom.n_cell[d] for d in range(3)]
    G = np.full([N[d] + (2 * grow if d < dim else 0) for d in range(3)], np.nan)
    for ib, (lo, hi) in enumerate(geom.boxes):
        a = mf[ib][..., comp]
        src, dst = [], []
        for d in range(3):
            if d >= dim:
                src.append(slice(None)); dst.append(slice(None)); continue
            g0 = grow if lo[d] == geom.dlo[d] else 0
            g1 = grow if hi[d] == geom.dhi[d] else 0
            src.append(slice(ng - g0, a.shape[d] - ng + g1))
            dst.append(slice(lo[d] + grow - g0, hi[d] + 1 + grow + g1))
        G[tuple(dst)] = a[tuple(src)]
    return G


def _scatter(geom, G, mf, ng, comp):
    for ib, (lo, hi) in enumerate(geom.boxes):
        sl = tuple(slice(lo[d], hi[d] + 1) if d < geom.dim else slice(None) for d in range(3))
        valid(geom, mf[ib], ib, ng)[..., comp] = G[sl]


def helm_ell_bc(geom, comp_is_vel, comp):
    """ell_bc_level_build (define_bc_tower.f90:254-340) for a velocity component / a scalar on the domain faces"""
    ell = np.zeros((3, 2), dtype=np.int32)
    ELL_PER, ELL_DIR, ELL_NEU = -1, 1, 2
    for d in range(geom.dim):
        for s in range(2):
            p = int(geom.phys_bc[d, s])
            if p == PERIODIC:
                ell[d, s] = ELL_PER
            elif p in (SLIP_WALL, SYMMETRY):
                ell[d, s] = ELL_DIR if (comp_is_vel and comp == d) else ELL_NEU
            elif p == NO_SLIP_WALL:
                ell[d, s] = ELL_DIR if comp_is_vel else ELL_NEU
            elif p == INLET:
                ell[d, s] = ELL_DIR
            elif p == OUTLET:
                ell[d, s] = ELL_NEU
            else:
                raise ValueError(p)
    return ell


def helm_rhs(geom, mf, ng, comp, comp_is_vel, rho, lap, mac_rhs, mu, diffusion_type, fold=True):
    """right-hand side of one Helmholtz solve on the whole domain: mkrhs_2d/3d (viscsolve.f90:193-299 for a velocity component, :464-513 for a
    scalar) in the reference's operation order, and -- fold=True -- the Dirichlet data: a ghost cell of an EXT_DIR face holds the boundary value,
    which contributes 8/3 mu phi_b / h^2 (the inhomogeneous part of the stencil_order-2 boundary stencil).
    -> (rh, field with one ghost layer, alpha, elliptic boundary types)"""
    dim = geom.dim
    N = [geom.n_cell[d] for d in range(3)]
    ug = _gather(geom, mf, ng, comp, grow=1)
    V = tuple(slice(1, N[d] + 1) if d < dim else slice(None) for d in range(3))
    u = ug[V]
    alpha = _gather(geom, rho, 1, 0) if comp_is_vel else np.ones(N)
    rh = u * alpha if comp_is_vel else u.copy()
    if diffusion_type == 1:
        rh = rh + mu * _gather(geom, lap, 0, comp)
    if comp_is_vel:
        visc_mu_dt = 2.0 * mu if diffusion_type == 1 else mu
        mg = _gather(geom, mac_rhs, 1, 0, grow=1)
        hi = [slice(None)] * 3; lo = [slice(None)] * 3
        for d in range(dim):
            hi[d] = slice(2, N[d] + 2) if d == comp else slice(1, N[d] + 1)
            lo[d] = slice(0, N[d]) if d == comp else slice(1, N[d] + 1)
        rh = rh + (1.0 / 3.0) * visc_mu_dt * (mg[tuple(hi)] - mg[tuple(lo)]) / geom.dx[comp]
    ell = helm_ell_bc(geom, comp_is_vel, comp)
    for d in range(dim):
        h2 = 1.0 / geom.dx[d] ** 2
        for s in range(2):
            if ell[d, s] != 1 or not fold:
                continue
            cell = [slice(None)] * 3; ghost = list(V)
            cell[d] = 0 if s == 0 else N[d] - 1
            ghost[d] = 0 if s == 0 else N[d] + 1
            rh[tuple(cell)] = rh[tuple(cell)] + ((8.0 / 3.0) * mu * h2) * ug[tuple(ghost)]
    return rh, ug, alpha, ell


def _helm_component(geom, params, mf, ng, comp, comp_is_vel, rho, lap, mac_rhs, mu, diffusion_type, rel_eps=1e-12):
    """one Helmholtz solve (alpha - mu div grad) phi = rh on component comp of mf, initial guess = the current field"""
    dim = geom.dim
    N = [geom.n_cell[d] for d in range(3)]
    rh, ug, alpha, ell = helm_rhs(geom, mf, ng, comp, comp_is_vel, rho, lap, mac_rhs, mu, diffusion_type)
    u = ug[tuple(slice(1, N[d] + 1) if d < dim else slice(None) for d in range(3))]
    nn = (C.c_int * 3)(*N)
    hh = (C.c_double * 3)(*[geom.dx[d] if d < dim else 1.0 for d in range(3)])
    eb = (C.c_int * 6)(*[int(x) for x in ell.ravel()])
    b = [np.full([N[t] + (1 if t == d else 0) for t in range(3)], mu, order="F") for d in range(dim)]
    phi = np.zeros([N[d] + 2 if d < dim else 1 for d in range(3)], order="F")
    F = lambda a: np.asfortranarray(a, dtype=np.float64)
    rhF, alF, u0F = F(rh), F(alpha), F(u)
    res = C.c_double(0.0)
    f = lib().orc_mg_solve_ex
    f.restype = C.c_int
    cyc = f(C.c_int(dim), nn, hh, eb, _dp(rhF), _dp(b[0]), _dp(b[1]), _dp(b[2]) if dim == 3 else None, _dp(alF), _dp(u0F), _dp(phi),
            C.c_double(rel_eps), C.c_int(params.mg_max_cycles), C.c_int(params.mg_nu1), C.c_int(params.mg_nu2), C.c_double(1e-3),
            C.c_int(params.mg_verbose), C.byref(res))
    PV = tuple(slice(1, N[d] + 1) if d < dim else slice(None) for d in range(3))
    _scatter(geom, phi[PV], mf, ng, comp)
    return cyc, res.value


def visc_solve(geom, params, unew, lapu, rho, mac_rhs, mu, diffusion_type=1):
    """viscsolve.f90:19-146: unew (ng 3, ghost cells filled) is solved component by component in place, then refilled (:105)"""
    tot, rmax = 0, 0.0
    for d in range(geom.dim):
        cyc, res = _helm_component(geom, params, unew, 3, d, True, rho, lapu, mac_rhs, mu, diffusion_type)
        tot += cyc; rmax = max(rmax, res)
    fill_and_physbc(geom, params, unew, 3, geom.dim, 0, 0, geom.dim)
    return tot, rmax


def diff_scalar_solve(geom, params, snew, laps, mu, icomp, diffusion_type=2):
    """viscsolve.f90:310-423 on component icomp (0-based) of snew, then fill_boundary + physbc of the scalars (:379-382)"""
    cyc, res = _helm_component(geom, params, snew, 3, icomp, False, None, laps, None, mu, diffusion_type)
    fill_and_physbc(geom, params, snew, 3, params.nscal, 0, geom.dim, params.nscal)
    return cyc, res


def project_with_phi(geom, params, umac_pred, rho, ncomp_s, phi):
    """mk_mac_coeffs + mkumac + fill_boundary(umac) for a GIVEN phi (whose ghost cells get filled here); returns umac."""
    dim = geom.dim
    beta = [mf_alloc(geom, 0, 1, d) for d in range(dim)]
    mk_mac_coeffs(geom, rho, 3, beta)
    ph = [a.copy(order='F') for a in phi]
    fill_boundary(geom, ph, 1, 1)
    um = [[a.copy(order='F') for a in umac_pred[d]] for d in range(dim)]
    dx = (C.c_double * 3)(*geom.dx)
    for ib, (lo, hi) in enumerate(geom.boxes):
        _, ell = _box_bc(geom, ib)
        lib().orc_mkumac(_dp(um[0][ib]), _dp(um[1][ib]), _dp(um[2][ib]) if dim == 3 else None, C.c_int(1), _dp(ph[ib]), C.c_int(1),
                         _dp(beta[0][ib]), _dp(beta[1][ib]), _dp(beta[2][ib]) if dim == 3 else None, C.c_int(0),
                         _ia(lo), _ia(hi), C.c_int(dim), dx, ell)
    for d in range(dim):
        fill_boundary(geom, um[d], 1, 1, face_dir=d)
    return um
