"""
oracle/ref.py -- TEST INFRASTRUCTURE ONLY.

ctypes harness around oracle/_ref/libref.so: the reference's OWN per-box Fortran routines, mechanically transpiled to C
by oracle/f2c.py from the sources under /root/reference (see that file's header for the list and the semantics kept).
It exists to PIN the hand-written oracle (oracle/orc_*.c): tests/test_ref_pin.py feeds both the same inputs and demands
bit-identical outputs, and tests/golden/make_golden.py stores small input/output vectors generated with it, so the pin
also holds where /root/reference is not mounted (the GPU box, the driver's CPU test run).

What is reference arithmetic here and what is not:
  * every `call(name, ...)` runs transpiled reference code (slope/velpred/mkflux/update/physbc/mkforce/macproject glue);
  * the multifab-level loops below mirror the reference drivers' argument passing (file:line cited per function);
  * `multifab_fill_boundary` / `ml_restrict_and_fill` live in FBoxLib, which is absent: the box<->box copy comes from
    oracle.fill_boundary (our restatement), followed by the reference's own physbc_2d/3d;
  * the multigrid (F_MG) is absent: nothing here solves for phi.
"""
import ctypes as C
import os
import numpy as np

from . import oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libref.so")
_SIG = os.path.join(_HERE, "_ref", "signatures.txt")

BC_PER, BC_INT, BC_DIR, BC_NEU = -1, 0, 1, 2


class _FA(C.Structure):
    _fields_ = [("p", C.c_void_p), ("ext", C.c_long * 4), ("st", C.c_long * 4)]


_lib, _sigs = None, None


def available():
    return os.path.exists(_SO) and os.path.exists(_SIG)


def lib():
    global _lib, _sigs
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libref.so missing: run `make -C oracle ref` where /root/reference is mounted")
        _lib = C.CDLL(_SO)
        _sigs = {}
        for ln in open(_SIG):
            parts = ln.split()
            _sigs[parts[0]] = (parts[1], [tuple(p.split(":")) for p in parts[2:]])
    return _lib


def where(name):
    """reference file:line of a transpiled routine"""
    lib()
    return _sigs[name][0]


def _desc(a, typ, rank, keep):
    want = np.float64 if typ == "real" else np.int32
    if not isinstance(a, np.ndarray):
        a = np.asarray(a, dtype=want)
        if a.ndim == 0:
            a = a.reshape(1)
        a = np.asfortranarray(a)
    if a.dtype != want:
        raise TypeError("array dtype %s, routine wants %s" % (a.dtype, want))
    if a.ndim != rank:
        raise ValueError("array rank %d, routine wants %d" % (a.ndim, rank))
    d = _FA()
    d.p = a.ctypes.data
    for i in range(rank):
        d.ext[i] = a.shape[i]
        d.st[i] = a.strides[i] // a.itemsize
    keep.append(a)
    return d


def call(name, *args):
    """Run the transpiled reference subroutine `name` with positional arguments in the Fortran order."""
    L = lib()
    _, sig = _sigs[name]
    if len(args) != len(sig):
        raise TypeError("%s takes %d arguments (%s), got %d" % (name, len(sig), " ".join(s[0] for s in sig), len(args)))
    keep, cargs, ios = [], [], []
    for a, (an, typ, rank) in zip(args, sig):
        rank = int(rank)
        if rank:
            cargs.append(C.byref(_desc(a, typ, rank, keep)))
        elif typ == "realio":                   # intent(inout) / intent(out) real scalar: by address, value returned
            ios.append(C.c_double(float(a)))
            cargs.append(C.byref(ios[-1]))
        elif typ == "real":
            cargs.append(C.c_double(float(a)))
        else:
            cargs.append(C.c_int(int(a)))
    C.c_int.in_dll(L, "ref_error_flag").value = 0
    getattr(L, "ref_" + name)(*cargs)
    if C.c_int.in_dll(L, "ref_error_flag").value:
        raise RuntimeError("reference routine %s called bl_error" % name)
    return [x.value for x in ios]


def set_probin(params):
    """probin_module values the per-box routines read (src/_parameters, probin.template:21-23)."""
    L = lib()
    for k, v in (("slope_order", params.slope_order), ("use_minion", params.use_minion), ("boussinesq", params.boussinesq),
                 ("nscal", params.nscal), ("diffusion_type", getattr(params, "diffusion_type", 1))):
        C.c_int.in_dll(L, k).value = int(v)
    C.c_double.in_dll(L, "visc_coef").value = params.visc_coef
    C.c_double.in_dll(L, "diff_coef").value = params.diff_coef
    for c, nm in enumerate(("u_bc", "v_bc", "w_bc", "rho_bc", "trac_bc")):
        arr = (C.c_double * 6).in_dll(L, nm)
        flat = np.asarray(params.bcval[c], dtype=float).reshape(3, 2).ravel(order="F")     # u_bc(dir, side), column major
        for i in range(6):
            arr[i] = flat[i]


# ---------------------------------------------------------------------------------------------------------------------
# BC tables, restated independently of orc_driver.c from define_bc_tower.f90
# ---------------------------------------------------------------------------------------------------------------------
def box_phys_bc(geom, ib):
    """phys_bc_level_array(i,:,:), define_bc_tower.f90:129-156"""
    dm = geom.dim
    lo, hi = geom.boxes[ib]
    pb = np.full((dm, 2), O.INTERIOR, dtype=np.int32, order="F")
    for d in range(dm):
        if lo[d] == geom.dlo[d]:
            pb[d, 0] = geom.phys_bc[d, 0]
        if hi[d] == geom.dhi[d]:
            pb[d, 1] = geom.phys_bc[d, 1]
    return pb


def adv_bc(pb, nscal):
    """adv_bc_level_array(i,:,:,:), define_bc_tower.f90:158-252  -> (dm, 2, dm+nscal+2)"""
    dm = pb.shape[0]
    adv = np.full((dm, 2, dm + nscal + 2), O.INTERIOR, dtype=np.int32, order="F")
    press, extrap = dm + nscal, dm + nscal + 1
    for d in range(dm):
        for s in range(2):
            p = pb[d, s]
            if p == O.SLIP_WALL:
                adv[d, s, :dm] = O.HOEXTRAP
                adv[d, s, d] = O.EXT_DIR
                adv[d, s, dm:dm + nscal] = O.HOEXTRAP
                adv[d, s, press] = O.FOEXTRAP
                adv[d, s, extrap] = O.FOEXTRAP
            elif p == O.NO_SLIP_WALL:
                adv[d, s, :dm] = O.EXT_DIR
                adv[d, s, dm:dm + nscal] = O.HOEXTRAP
                adv[d, s, press] = O.FOEXTRAP
                adv[d, s, extrap] = O.FOEXTRAP
            elif p == O.INLET:
                adv[d, s, :dm] = O.EXT_DIR
                adv[d, s, dm:dm + nscal] = O.EXT_DIR
                adv[d, s, press] = O.FOEXTRAP
                adv[d, s, extrap] = O.FOEXTRAP
            elif p == O.OUTLET:
                adv[d, s, :dm] = O.FOEXTRAP
                adv[d, s, dm:dm + nscal] = O.FOEXTRAP
                adv[d, s, press] = O.EXT_DIR
                adv[d, s, extrap] = O.FOEXTRAP
            elif p == O.SYMMETRY:
                adv[d, s, :dm] = O.REFLECT_EVEN
                adv[d, s, d] = O.REFLECT_ODD
                adv[d, s, dm:dm + nscal] = O.REFLECT_EVEN
                adv[d, s, press] = O.EXT_DIR
                adv[d, s, extrap] = O.REFLECT_EVEN
    return adv


def ell_bc_press(pb):
    """ell_bc_level_array(i,:,:,press_comp), define_bc_tower.f90:254-340"""
    dm = pb.shape[0]
    ell = np.full((dm, 2), BC_INT, dtype=np.int32, order="F")
    for d in range(dm):
        for s in range(2):
            p = pb[d, s]
            if p in (O.SLIP_WALL, O.NO_SLIP_WALL, O.INLET, O.SYMMETRY):
                ell[d, s] = BC_NEU
            elif p == O.OUTLET:
                ell[d, s] = BC_DIR
            elif p == O.PERIODIC:
                ell[d, s] = BC_PER
    return ell


# ---------------------------------------------------------------------------------------------------------------------
# multifab-level stages: the reference drivers' fab loops with the reference's argument passing
# ---------------------------------------------------------------------------------------------------------------------
def _lohi(geom, ib):
    lo, hi = geom.boxes[ib]
    dm = geom.dim
    return np.asarray(lo[:dm], dtype=np.int32), np.asarray(hi[:dm], dtype=np.int32)


def _c(a, dim):
    """4-D box array -> what the driver passes: ap(:,:,1,:) in 2-D, ap(:,:,:,:) in 3-D"""
    return a[:, :, 0, :] if dim == 2 else a


def _c1(a, dim, comp=0):
    """single component: ap(:,:,1,c) / ap(:,:,:,c)"""
    return a[:, :, 0, comp] if dim == 2 else a[:, :, :, comp]


def _dx(geom):
    return np.asarray(geom.dx[:geom.dim], dtype=np.float64)


def physbc_mf(geom, params, mf, ng, start_scomp, start_bccomp, num_comp):
    """multifab_physbc, multifab_physbc.f90:17-62 (components 1-based like the reference)."""
    if ng == 0:
        return
    set_probin(params)
    dm = geom.dim
    for ib in range(geom.nboxes):
        lo, hi = _lohi(geom, ib)
        adv = adv_bc(box_phys_bc(geom, ib), params.nscal)
        for scomp in range(start_scomp, start_scomp + num_comp):
            bccomp = start_bccomp + scomp - start_scomp
            call("physbc_%dd" % dm, _c1(mf[ib], dm, scomp - 1), lo, hi, ng, np.asfortranarray(adv[:, :, bccomp - 1]), bccomp)


def restrict_and_fill(geom, params, mf, ng, ncomp, icomp, bcomp, nc, same_boundary=False):
    """ml_restrict_and_fill for nlevs == 1 (FBoxLib, absent): fill_boundary, then multifab_physbc per component."""
    O.fill_boundary(geom, mf, ng, ncomp)
    if same_boundary:
        for c in range(nc):
            physbc_mf(geom, params, mf, ng, icomp + c, bcomp, 1)
    else:
        physbc_mf(geom, params, mf, ng, icomp, bcomp, nc)


def mkvelforce(geom, params, vel_force, ext, gp, s, ng_s, lapu, visc_fac):
    """mkforce.f90:18-80"""
    set_probin(params)
    dm = geom.dim
    for ib in range(geom.nboxes):
        lo, hi = _lohi(geom, ib)
        vel_force[ib][...] = 0.0
        sb = s[ib]
        if sb.shape[3] < dm:      # rhohalf: the reference builds it with dm components, all zero but the first (advance_timestep.f90:70,73)
            sb = np.concatenate([sb, np.zeros(sb.shape[:3] + (dm - sb.shape[3],))], axis=3).copy(order="F")
        call("mkvelforce_%dd" % dm, _c(vel_force[ib], dm), _c(ext[ib], dm), _c(gp[ib], dm), _c(sb, dm), _c(lapu[ib], dm),
             1, 1, 1, ng_s, 0, visc_fac, lo, hi)
    extrap_comp = dm + params.nscal + 2
    restrict_and_fill(geom, params, vel_force, 1, dm, 1, extrap_comp, dm, same_boundary=True)


def mkscalforce(geom, params, scal_force, ext, laps, diff_fac):
    """mkforce.f90:238-288"""
    set_probin(params)
    dm = geom.dim
    for ib in range(geom.nboxes):
        lo, hi = _lohi(geom, ib)
        scal_force[ib][...] = 0.0
        call("mkscalforce_%dd" % dm, _c(scal_force[ib], dm), _c(ext[ib], dm), _c(laps[ib], dm), 1, 1, 0, diff_fac, lo, hi)
    extrap_comp = dm + params.nscal + 2
    restrict_and_fill(geom, params, scal_force, 1, params.nscal, 1, extrap_comp, params.nscal, same_boundary=True)


def velpred(geom, params, u, umac, force, dt, debug=False):
    """velpred.f90:16-123 (single level): per-fab kernel, then fill_boundary(umac(d))."""
    set_probin(params)
    dm = geom.dim
    name = ("velpred_debug_%dd" if debug else "velpred_%dd") % dm
    for ib in range(geom.nboxes):
        lo, hi = _lohi(geom, ib)
        pb = box_phys_bc(geom, ib)
        adv = adv_bc(pb, params.nscal)
        um = [_c1(umac[d][ib], dm) for d in range(dm)]
        call(name, _c(u[ib], dm), *um, _c(force[ib], dm), lo, hi, _dx(geom), dt, pb, adv, 3, 1, 1)
    for d in range(dm):
        O.fill_boundary(geom, umac[d], 1, 1, face_dir=d)


def mkflux(geom, params, sold, ncomp, sedge, flux, umac, force, mac_rhs, dt, is_vel, is_cons, debug=False):
    """mkflux.f90:16-150 (single level)"""
    set_probin(params)
    dm = geom.dim
    name = ("mkflux_debug_%dd" if debug else "mkflux_%dd") % dm
    bccomp = 1 if is_vel else dm + 1
    ic = np.asarray([int(x) for x in is_cons[:ncomp]], dtype=np.int32)
    for ib in range(geom.nboxes):
        lo, hi = _lohi(geom, ib)
        pb = box_phys_bc(geom, ib)
        adv = np.asfortranarray(adv_bc(pb, params.nscal)[:, :, bccomp - 1:bccomp - 1 + ncomp])
        call(name, _c(sold[ib], dm), *[_c(sedge[d][ib], dm) for d in range(dm)], *[_c(flux[d][ib], dm) for d in range(dm)],
             *[_c1(umac[d][ib], dm) for d in range(dm)], _c(force[ib], dm), _c1(mac_rhs[ib], dm),
             lo, hi, _dx(geom), dt, int(is_vel), pb, adv, 3, 0, 0, 1, 1, 1, ic)


def update(geom, params, sold, ncomp, umac, sedge, flux, force, snew, dt, is_vel, is_cons):
    """update.f90:16-111 (single level)"""
    set_probin(params)
    dm = geom.dim
    ic = np.asarray([int(x) for x in is_cons[:ncomp]], dtype=np.int32)
    for ib in range(geom.nboxes):
        lo, hi = _lohi(geom, ib)
        call("update_%dd" % dm, _c(sold[ib], dm), *[_c1(umac[d][ib], dm) for d in range(dm)],
             *[_c(sedge[d][ib], dm) for d in range(dm)], *[_c(flux[d][ib], dm) for d in range(dm)],
             _c(force[ib], dm), _c(snew[ib], dm), lo, hi, 3, 1, 0, 0, 1, _dx(geom), dt, int(is_vel), ic)
    restrict_and_fill(geom, params, snew, 3, ncomp, 1, 1 if is_vel else dm + 1, ncomp)


def make_at_halftime(geom, params, rhohalf, sold, snew):
    """make_at_halftime.f90:18-71 with in_comp = out_comp = 1 (advance_timestep.f90:114)"""
    set_probin(params)
    dm = geom.dim
    for ib in range(geom.nboxes):
        lo, hi = _lohi(geom, ib)
        call("make_at_halftime_%dd" % dm, _c1(rhohalf[ib], dm), _c1(sold[ib], dm), _c1(snew[ib], dm), lo, hi, 1, 3)
    restrict_and_fill(geom, params, rhohalf, 1, 1, 1, dm + 1, 1)


def estdt(geom, u, ng_u, s, ng_s, gp, ng_g, ext_vel_force, ng_f, dtold=-1.0, cflfac=0.5, max_dt_growth=1.1):
    """estdt.f90:15-87 around the reference's own estdt_2d / estdt_3d (:89-181)"""
    dm = geom.dim
    dt = dt_start = 1.e20
    for ib in range(geom.nboxes):
        lo, hi = _lohi(geom, ib)
        (dt_grid,) = call("estdt_%dd" % dm, _c(u[ib], dm), ng_u, _c1(s[ib], dm), ng_s, _c(gp[ib], dm), ng_g, _c(ext_vel_force[ib], dm), ng_f,
                          lo, hi, _dx(geom), 1.e20)
        dt = min(dt_grid, dt)
    if dt == dt_start:
        dt = min(_dx(geom)[:dm])
    dt = dt * cflfac
    if dtold > 0.0:
        dt = min(dt, max_dt_growth * dtold)
    return dt


def visc_mkrhs(geom, unew, lapu, rho, mac_rhs, mu, comp, diffusion_type):
    """visc_solve's internal mkrhs (viscsolve.f90:147-191) around the reference's own mkrhs_2d / mkrhs_3d (:193-299): per box rh (no ghost
    cells) and phi (one ghost layer) for velocity component comp (0-based)"""
    dm = geom.dim
    C.c_int.in_dll(lib(), "diffusion_type").value = int(diffusion_type)
    rh, phi = O.mf_alloc(geom, 0, 1), O.mf_alloc(geom, 1, 1)
    for ib in range(geom.nboxes):
        call("visc_mkrhs_%dd" % dm, _c1(rh[ib], dm), _c1(unew[ib], dm, comp), _c1(lapu[ib], dm, comp), _c1(rho[ib], dm), _c1(phi[ib], dm),
             _c1(mac_rhs[ib], dm), mu, _dx(geom), 3, 1, 1, comp + 1)
    return rh, phi


def scal_mkrhs(geom, snew, laps, mu, comp, diffusion_type):
    """diff_scalar_solve's internal mkrhs (viscsolve.f90:426-462) around the reference's own mkrhs_2d / mkrhs_3d (:464-513)"""
    dm = geom.dim
    C.c_int.in_dll(lib(), "diffusion_type").value = int(diffusion_type)
    rh, phi = O.mf_alloc(geom, 0, 1), O.mf_alloc(geom, 1, 1)
    for ib in range(geom.nboxes):
        call("scal_mkrhs_%dd" % dm, _c1(rh[ib], dm), _c1(snew[ib], dm, comp), _c1(laps[ib], dm, comp), _c1(phi[ib], dm), mu, 3)
    return rh, phi


def divumac(geom, umac, mac_rhs, rh):
    """macproject.f90:137-225: rh = mac_rhs - div(umac)   (mult_mult_s(-1) then plus_plus)"""
    dm = geom.dim
    for ib in range(geom.nboxes):
        lo, hi = _lohi(geom, ib)
        call("divumac_%dd" % dm, *[_c1(umac[d][ib], dm) for d in range(dm)], 1, _c1(rh[ib], dm), 0, _dx(geom), lo, hi)
        rh[ib][...] = rh[ib] * (-1.0)
        v = O.valid(geom, mac_rhs[ib], ib, 1)
        rh[ib][...] = rh[ib] + v


def mk_mac_coeffs(geom, rho, ng_r, beta):
    """macproject.f90:280-336"""
    dm = geom.dim
    for ib in range(geom.nboxes):
        lo, hi = _lohi(geom, ib)
        call("mk_mac_coeffs_%dd" % dm, *[_c1(beta[d][ib], dm) for d in range(dm)], 0, _c1(rho[ib], dm), ng_r, lo, hi)


def mkumac(geom, params, umac, phi, beta, fine_flx):
    """macproject.f90:403-505.  fine_flx[d][side][ib]: arrays shaped like the box with extent 1 in direction d."""
    dm = geom.dim
    for ib in range(geom.nboxes):
        lo, hi = _lohi(geom, ib)
        ell = ell_bc_press(box_phys_bc(geom, ib))
        fl = []
        for d in range(dm):
            for side in range(2):
                fl.append(_c1(fine_flx[d][side][ib], dm))
        call("mkumac_%dd" % dm, *[_c1(umac[d][ib], dm) for d in range(dm)], 1, _c1(phi[ib], dm), 1,
             *[_c1(beta[d][ib], dm) for d in range(dm)], 0, *fl, lo, hi, _dx(geom), ell)
    for d in range(dm):
        O.fill_boundary(geom, umac[d], 1, 1, face_dir=d)


# ---------------------------------------------------------------------------------------------------------------------
# stage-by-stage pin of the hand-written oracle against the transpiled reference
# ---------------------------------------------------------------------------------------------------------------------
def _cp(mf):
    return [a.copy(order="F") for a in mf]


def fine_flx_from_phi(geom, phi, beta):
    """What F_MG (absent) hands back in fine_flx: the stencil flux through each box's boundary faces, outward sign on
    the hi side (macproject.f90:608-609: umac(lo) -= lo_flx*dx, umac(hi+1) += hi_flx*dx).  phi ghost cells must be filled."""
    dm = geom.dim
    out = [[[], []] for _ in range(dm)]
    for ib in range(geom.nboxes):
        ell = ell_bc_press(box_phys_bc(geom, ib))
        p = phi[ib][..., 0]
        for d in range(dm):
            h = geom.dx[d]
            b = beta[d][ib][..., 0]
            n = b.shape[d] - 1                                   # cells in direction d
            v = [slice(1, -1) if q < dm and q != d else slice(None) for q in range(3)]   # valid transverse range of phi

            def at(idx):
                s = list(v)
                s[d] = slice(idx, idx + 1)
                return p[tuple(s)]

            def bt(idx):
                s = [slice(None)] * 3
                s[d] = slice(idx, idx + 1)
                return b[tuple(s)]
            # phi index: ghost = 0, first valid = 1, last valid = n, hi ghost = n+1
            if ell[d, 0] == BC_NEU:
                lo_f = np.zeros_like(bt(0))
            elif ell[d, 0] == BC_DIR:
                lo_f = bt(0) * (3.0 * at(1) - at(2) / 3.0) / h / h
            else:
                lo_f = bt(0) * (at(1) - at(0)) / h / h
            if ell[d, 1] == BC_NEU:
                hi_f = np.zeros_like(bt(n))
            elif ell[d, 1] == BC_DIR:
                hi_f = bt(n) * (3.0 * at(n) - at(n - 1) / 3.0) / h / h
            else:
                hi_f = -bt(n) * (at(n + 1) - at(n)) / h / h
            out[d][0].append(np.asfortranarray(lo_f)[..., None].copy(order="F"))
            out[d][1].append(np.asfortranarray(hi_f)[..., None].copy(order="F"))
    return out


def stagewise_from(geom, params, st, dt, o, debug=False):
    """Every stage of the path run with the TRANSPILED REFERENCE routines on the inputs the oracle's own stage saw
    (o = oracle.stagewise(...)); returns the reference outputs under the oracle's key names."""
    dim, nscal = geom.dim, params.nscal
    mf_alloc = O.mf_alloc
    lapu = st.get("lapu") or mf_alloc(geom, 0, dim)
    r = {}
    r["vel_force_1"] = mf_alloc(geom, 1, dim)
    mkvelforce(geom, params, r["vel_force_1"], st["ext_vel_force"], st["gp"], st["sold"], 3, lapu, 1.0)
    r["umac_pred"] = [mf_alloc(geom, 1, 1, d, val=1.0e20) for d in range(dim)]
    velpred(geom, params, st["uold"], r["umac_pred"], o["vel_force_1"], dt, debug=debug)
    # macproject glue on the oracle's predicted umac / phi
    mac_rhs = mf_alloc(geom, 1, 1)
    r["rh"] = mf_alloc(geom, 0, 1)
    divumac(geom, o["umac_pred"], mac_rhs, r["rh"])
    r["beta"] = [mf_alloc(geom, 0, 1, d) for d in range(dim)]
    mk_mac_coeffs(geom, st["sold"], 3, r["beta"])
    phi = _cp(o["phi"])
    O.fill_boundary(geom, phi, 1, 1)
    r["umac"] = [_cp(o["umac_pred"][d]) for d in range(dim)]
    mkumac(geom, params, r["umac"], phi, r["beta"], fine_flx_from_phi(geom, phi, r["beta"]))
    # scalars
    laps = mf_alloc(geom, 0, nscal)
    r["scal_force_1"] = mf_alloc(geom, 1, nscal)
    mkscalforce(geom, params, r["scal_force_1"], st["ext_scal_force"], laps, 1.0)
    r["sedge"] = [mf_alloc(geom, 0, nscal, d) for d in range(dim)]
    r["sflux"] = [mf_alloc(geom, 0, nscal, d) for d in range(dim)]
    divu = mf_alloc(geom, 1, 1)
    is_cons_s = [1] + [0] * (nscal - 1)
    mkflux(geom, params, st["sold"], nscal, r["sedge"], r["sflux"], o["umac"], o["scal_force_1"], divu, dt, False, is_cons_s,
           debug=debug)
    r["scal_force_2"] = mf_alloc(geom, 1, nscal)
    mkscalforce(geom, params, r["scal_force_2"], st["ext_scal_force"], laps, 0.0)
    r["snew"] = mf_alloc(geom, 3, nscal)
    update(geom, params, st["sold"], nscal, o["umac"], o["sedge"], o["sflux"], o["scal_force_2"], r["snew"], dt, False, is_cons_s)
    r["rhohalf"] = mf_alloc(geom, 1, 1)
    make_at_halftime(geom, params, r["rhohalf"], st["sold"], o["snew"])
    r["uedge"] = [mf_alloc(geom, 0, dim, d) for d in range(dim)]
    uflux = [mf_alloc(geom, 0, dim, d) for d in range(dim)]
    mkflux(geom, params, st["uold"], dim, r["uedge"], uflux, o["umac"], o["vel_force_1"], mac_rhs, dt, True, [0] * dim, debug=debug)
    r["vel_force_2"] = mf_alloc(geom, 1, dim)
    mkvelforce(geom, params, r["vel_force_2"], st["ext_vel_force"], st["gp"], o["rhohalf"], 1, lapu, 0.0)
    r["unew"] = mf_alloc(geom, 3, dim)
    update(geom, params, st["uold"], dim, o["umac"], o["uedge"], uflux, o["vel_force_2"], r["unew"], dt, True, [0] * dim)
    return r


PIN_KEYS = ["vel_force_1", "umac_pred", "umac", "scal_force_1", "sedge", "sflux", "scal_force_2", "snew", "rhohalf", "uedge",
            "vel_force_2", "unew"]


def flatten(x):
    """multifab or list of multifabs -> flat list of arrays"""
    if isinstance(x, np.ndarray):
        return [x]
    res = []
    for y in x:
        res += flatten(y)
    return res


def max_diff(a, b):
    """(max |a-b|, max |b|, number of differing entries) over a (nested) multifab; NaNs count as differences"""
    md, mb, nd = 0.0, 0.0, 0
    for x, y in zip(flatten(a), flatten(b)):
        ne = ~((x == y) | (np.isnan(x) & np.isnan(y)))
        nd += int(ne.sum())
        if ne.any():
            d = np.abs(x[ne] - y[ne])
            md = max(md, float(np.nanmax(d)) if np.isfinite(d).any() else np.inf)
        fin = np.isfinite(y) & (np.abs(y) < 1e19)
        if fin.any():
            mb = max(mb, float(np.abs(y[fin]).max()))
    return md, mb, nd
