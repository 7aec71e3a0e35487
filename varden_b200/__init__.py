"""
varden_b200 -- B200-native (sm_100a, FP64) implementation of VARDEN's per-timestep
advection + MAC-projection hot path, behind the C ABI declared in include/vdn.h.

This package is only the ctypes binding used by tests/ and bench.py; the product is
varden_b200/libvdn.so (hand-written CUDA + the C++ host orchestration).  There is no CPU
fallback: importing works without a GPU (so the symbol table can be checked), but creating
a context without a CUDA device, or with the library missing, fails loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvdn.so")

# field ids: keep in sync with enum vdn_field in include/vdn.h
FIELDS = ["UOLD", "SOLD", "UNEW", "SNEW", "GP", "EXT_VEL_FORCE", "EXT_SCAL_FORCE", "LAPU",
          "UMAC_X", "UMAC_Y", "UMAC_Z", "MAC_RHS", "RHOHALF", "VEL_FORCE", "SCAL_FORCE",
          "RH", "PHI", "BETA_X", "BETA_Y", "BETA_Z", "SEDGE_X", "SEDGE_Y", "SEDGE_Z",
          "SFLUX_X", "SFLUX_Y", "SFLUX_Z", "UEDGE_X", "UEDGE_Y", "UEDGE_Z"]
F = {name: i for i, name in enumerate(FIELDS)}

ABI_SYMBOLS = [
    "vdn_params_default", "vdn_ctx_create", "vdn_ctx_destroy", "vdn_last_error", "vdn_ctx_set_comm",
    "vdn_nccl_unique_id", "vdn_comm_plan", "vdn_halo_plan", "vdn_halo_plan_ex", "vdn_comm_tune",
    "vdn_field_upload", "vdn_field_download", "vdn_field_setval", "vdn_sync", "vdn_get_stream",
    "vdn_fill_boundary", "vdn_fill_and_physbc", "vdn_mkvelforce", "vdn_mkscalforce", "vdn_velpred",
    "vdn_macproject", "vdn_mkflux", "vdn_update", "vdn_make_at_halftime", "vdn_advance", "vdn_advance_host",
    "vdn_divumac", "vdn_mk_mac_coeffs", "vdn_mac_solve", "vdn_mkumac",
    "vdn_prof_enable", "vdn_prof_count", "vdn_prof_get", "vdn_launch_count", "vdn_mg_tune", "vdn_device_count", "vdn_comm_bytes",
    "vdn_debug_counters", "vdn_estdt", "vdn_field_copy", "vdn_visc_solve", "vdn_diff_scalar_solve",
]


class VdnParams(C.Structure):
    _fields_ = [("nscal", C.c_int), ("slope_order", C.c_int), ("use_minion", C.c_int), ("boussinesq", C.c_int),
                ("stencil_order", C.c_int), ("mg_verbose", C.c_int), ("mg_nu1", C.c_int), ("mg_nu2", C.c_int),
                ("mg_max_cycles", C.c_int), ("mg_max_bottom_iter", C.c_int),
                ("mg_bottom_eps", C.c_double), ("visc_coef", C.c_double), ("diff_coef", C.c_double),
                ("bc_val", C.c_double * 30)]


class VdnHostState(C.Structure):
    """vdn_host_state: one pointer per local box and multifab (include/vdn.h)"""
    _fields_ = [(k, C.POINTER(C.POINTER(C.c_double))) for k in
                ("uold", "sold", "gp", "ext_vel_force", "ext_scal_force", "unew", "snew", "rhohalf", "lapu", "mac_rhs")]
    OPTIONAL = ("lapu", "mac_rhs")


class VdnError(RuntimeError):
    pass


_lib = None


def load_library():
    """Load libvdn.so; raise if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VdnError("varden_b200/libvdn.so is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a). The hot path has no CPU fallback.")
        # libvdn.so needs libnccl.so.2.  PyTorch ships a newer NCCL under the same SONAME than the system one; whichever is mapped
        # first serves both, so map PyTorch's first (a superset) -- otherwise a later `import torch` fails to resolve its symbols.
        try:
            import importlib.util
            spec = importlib.util.find_spec("nvidia")
            for base in (spec.submodule_search_locations if spec else []):
                cand = os.path.join(base, "nccl", "lib", "libnccl.so.2")
                if os.path.exists(cand):
                    C.CDLL(cand, mode=C.RTLD_GLOBAL)
                    break
        except Exception:
            pass
        _lib = C.CDLL(os.environ.get("VDN_LIB", LIB_PATH))      # VDN_LIB: a differently tuned build of the same library (kernel tuning runs)
        _lib.vdn_last_error.restype = C.c_char_p
        _lib.vdn_launch_count.restype = C.c_longlong
        _lib.vdn_comm_bytes.restype = C.c_longlong
        _lib.vdn_get_stream.restype = C.c_void_p
    return _lib


def default_params(**kw):
    lib = load_library()
    p = VdnParams()
    lib.vdn_params_default(C.byref(p))
    bc_val = kw.pop("bc_val", None)
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    if bc_val is not None:
        for i, v in enumerate(np.asarray(bc_val, dtype=float).ravel()):
            p.bc_val[i] = v
    return p


def _iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class Context:
    """Device mirror of one level's rank-local boxes (vdn_ctx).  Method names follow the reference procedures."""

    def __init__(self, dim, boxes, dom_lo, dom_hi, phys_bc, dx, params=None, device=0):
        self.lib = load_library()
        self.dim = dim
        self.boxes = [(list(lo), list(hi)) for lo, hi in boxes]
        self.params = params if params is not None else default_params()
        blo = np.ascontiguousarray([b[0] for b in self.boxes], dtype=np.int32).reshape(-1, 3)
        bhi = np.ascontiguousarray([b[1] for b in self.boxes], dtype=np.int32).reshape(-1, 3)
        dlo = np.ascontiguousarray(list(dom_lo) + [0] * (3 - len(dom_lo)), dtype=np.int32)
        dhi = np.ascontiguousarray(list(dom_hi) + [0] * (3 - len(dom_hi)), dtype=np.int32)
        pbc = np.zeros((3, 2), dtype=np.int32)
        pbc[:dim] = np.asarray(phys_bc, dtype=np.int32).reshape(-1, 2)[:dim]
        dxa = np.ascontiguousarray(list(dx)[:3] + [0.0] * (3 - len(dx)), dtype=np.float64)
        h = C.c_void_p()
        rc = self.lib.vdn_ctx_create(C.byref(self.params), dim, len(self.boxes), _iptr(blo), _iptr(bhi), _iptr(dlo), _iptr(dhi),
                                     _iptr(pbc), dxa.ctypes.data_as(C.POINTER(C.c_double)), device, C.byref(h))
        if rc != 0:
            raise VdnError(self.lib.vdn_last_error(None).decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.vdn_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise VdnError(self.lib.vdn_last_error(self.h).decode())

    # ---- path-boundary copies ----
    def upload(self, field, ibox, arr, ng, ncomp):
        """arr: numpy (order F, float64) or a pinned torch tensor's numpy view of one box incl. ghosts"""
        assert arr.dtype == np.float64 and arr.flags.f_contiguous
        self._chk(self.lib.vdn_field_upload(self.h, F[field], ibox, arr.ctypes.data_as(C.POINTER(C.c_double)), ng, ncomp))

    def upload_ptr(self, field, ibox, ptr, ng, ncomp):
        self._chk(self.lib.vdn_field_upload(self.h, F[field], ibox, C.cast(ptr, C.POINTER(C.c_double)), ng, ncomp))

    def download(self, field, ibox, arr, ng, ncomp):
        assert arr.dtype == np.float64 and arr.flags.f_contiguous
        self._chk(self.lib.vdn_field_download(self.h, F[field], ibox, arr.ctypes.data_as(C.POINTER(C.c_double)), ng, ncomp))

    def download_ptr(self, field, ibox, ptr, ng, ncomp):
        self._chk(self.lib.vdn_field_download(self.h, F[field], ibox, C.cast(ptr, C.POINTER(C.c_double)), ng, ncomp))

    def upload_mf(self, field, mf, ng, ncomp):
        for ib, a in enumerate(mf):
            self.upload(field, ib, a, ng, ncomp)

    def download_mf(self, field, mf, ng, ncomp):
        for ib, a in enumerate(mf):
            self.download(field, ib, a, ng, ncomp)

    def setval(self, field, val):
        self._chk(self.lib.vdn_field_setval(self.h, F[field], C.c_double(val)))

    def sync(self):
        self._chk(self.lib.vdn_sync(self.h))

    def stream_ptr(self):
        return int(self.lib.vdn_get_stream(self.h))

    # ---- stage calls (reference procedure names) ----
    def fill_boundary(self, field):
        self._chk(self.lib.vdn_fill_boundary(self.h, F[field]))

    def fill_and_physbc(self, field, bccomp, same_boundary=False):
        self._chk(self.lib.vdn_fill_and_physbc(self.h, F[field], bccomp, int(same_boundary)))

    def mkvelforce(self, rho_field="SOLD", visc_fac=1.0):
        self._chk(self.lib.vdn_mkvelforce(self.h, F[rho_field], C.c_double(visc_fac)))

    def mkscalforce(self, diff_fac=1.0):
        self._chk(self.lib.vdn_mkscalforce(self.h, C.c_double(diff_fac)))

    def velpred(self, dt):
        self._chk(self.lib.vdn_velpred(self.h, C.c_double(dt)))

    def mkflux(self, is_vel, dt):
        self._chk(self.lib.vdn_mkflux(self.h, int(is_vel), C.c_double(dt)))

    def update(self, is_vel, dt):
        self._chk(self.lib.vdn_update(self.h, int(is_vel), C.c_double(dt)))

    def make_at_halftime(self):
        self._chk(self.lib.vdn_make_at_halftime(self.h))

    def divumac(self, want_norm=True):
        v = C.c_double(0.0)
        self._chk(self.lib.vdn_divumac(self.h, C.byref(v) if want_norm else None))
        return v.value

    def mk_mac_coeffs(self):
        self._chk(self.lib.vdn_mk_mac_coeffs(self.h))

    def mkumac(self):
        self._chk(self.lib.vdn_mkumac(self.h))

    def mac_solve(self, rel_eps=1e-10, abs_eps=-1.0):
        n, r = C.c_int(0), C.c_double(0.0)
        self._chk(self.lib.vdn_mac_solve(self.h, C.c_double(rel_eps), C.c_double(abs_eps), C.byref(n), C.byref(r)))
        return n.value, r.value

    def estdt(self, dtold=-1.0, cflfac=0.5, max_dt_growth=1.1):
        """estdt.f90:15-87 on the resident UOLD / SOLD / GP / EXT_VEL_FORCE (probin defaults cflfac 0.5, max_dt_growth 1.1)"""
        dt = C.c_double(0.0)
        self._chk(self.lib.vdn_estdt(self.h, C.c_double(dtold), C.c_double(cflfac), C.c_double(max_dt_growth), C.byref(dt)))
        return dt.value

    def field_copy(self, dst, src):
        """dst <- src on the device, ghost cells included (varden.f90:321-324)"""
        self._chk(self.lib.vdn_field_copy(self.h, F[dst], F[src]))

    def visc_solve(self, mu, diffusion_type=1):
        """viscsolve.f90:19 on UNEW (alpha = RHOHALF, LAPU, MAC_RHS resident) -> (V-cycles summed over the components, largest rel. residual)"""
        n, r = C.c_int(0), C.c_double(0.0)
        self._chk(self.lib.vdn_visc_solve(self.h, C.c_double(mu), int(diffusion_type), C.byref(n), C.byref(r)))
        return n.value, r.value

    def diff_scalar_solve(self, mu, icomp, diffusion_type=2):
        """viscsolve.f90:310 on component icomp (0-based) of SNEW"""
        n, r = C.c_int(0), C.c_double(0.0)
        self._chk(self.lib.vdn_diff_scalar_solve(self.h, C.c_double(mu), int(icomp), int(diffusion_type), C.byref(n), C.byref(r)))
        return n.value, r.value

    def comm_tune(self, mode):
        """measurement hook, before set_comm: transport of the ghost exchanges -- 0 peer memory with the fused sweeps pushing their boundary results
        (default), 1 NCCL, 2 peer memory with a pull kernel before every sweep, 3 peer memory with a push kernel after every sweep"""
        self._chk(self.lib.vdn_comm_tune(self.h, int(mode)))

    def debug_counters(self):
        """measurement hook: flag-wait accounting of the fused smoother per kernel family {name: (ms waited, longest wait in us, waits)};
        the first call switches it on, every call resets it"""
        buf = (C.c_ulonglong * 32)()
        self._chk(self.lib.vdn_debug_counters(self.h, buf))
        names = ["mg_wave_smooth_l0", "mg_wave_down_l0", "mg_wave_pro_l0", "mg_wave_up_l0", "mg_wave_coarse"]
        return {n: (buf[4 * i] / 1e6, buf[4 * i + 1] / 1e3, int(buf[4 * i + 2])) for i, n in enumerate(names)}

    def mg_tune(self, fuse_min=128, tile=-1):
        """test hook: smallest level the fused smoother runs on, forced tile shape (-1: measured defaults)"""
        self._chk(self.lib.vdn_mg_tune(self.h, int(fuse_min), int(tile)))

    def macproject(self, rel_eps=-1.0, abs_eps=-1.0):
        n, r = C.c_int(0), C.c_double(0.0)
        self._chk(self.lib.vdn_macproject(self.h, C.c_double(rel_eps), C.c_double(abs_eps), C.byref(n), C.byref(r)))
        return n.value, r.value

    def advance(self, dt, mac_rel_eps=-1.0):
        """advance_timestep.f90:95-124 on the device; returns (V-cycles, final relative residual)."""
        n, r = C.c_int(0), C.c_double(0.0)
        self._chk(self.lib.vdn_advance(self.h, C.c_double(dt), C.c_double(mac_rel_eps), C.byref(n), C.byref(r)))
        return n.value, r.value

    def host_state(self, **mfs):
        """build a vdn_host_state from lists of per-box numpy arrays (order F, float64; page-locked for overlapping copies)"""
        hs = VdnHostState()
        keep = []
        for k, _ in VdnHostState._fields_:
            if k in VdnHostState.OPTIONAL and mfs.get(k) is None:
                continue                                     # NULL: not handed over
            mf = mfs[k]
            assert len(mf) == len(self.boxes) and all(a.dtype == np.float64 and a.flags.f_contiguous for a in mf)
            arr = (C.POINTER(C.c_double) * len(mf))(*[a.ctypes.data_as(C.POINTER(C.c_double)) for a in mf])
            keep.append(arr)
            setattr(hs, k, C.cast(arr, C.POINTER(C.POINTER(C.c_double))))
        hs._keep = (keep, mfs)
        return hs

    def advance_host(self, dt, hs, mac_rel_eps=-1.0):
        """the same pass from / to HOST multifabs (vdn_advance_host): pipelined H2D -> stages -> D2H"""
        n, r = C.c_int(0), C.c_double(0.0)
        self._chk(self.lib.vdn_advance_host(self.h, C.c_double(dt), C.c_double(mac_rel_eps), C.byref(hs), C.byref(n), C.byref(r)))
        return n.value, r.value

    # ---- measurement ----
    def prof_enable(self, on=True):
        self._chk(self.lib.vdn_prof_enable(self.h, int(on)))

    def prof_report(self):
        out = {}
        n = self.lib.vdn_prof_count(self.h)
        for i in range(n):
            name = C.create_string_buffer(64)
            la, ms, by = C.c_longlong(0), C.c_double(0.0), C.c_double(0.0)
            self.lib.vdn_prof_get(self.h, i, name, C.byref(la), C.byref(ms), C.byref(by))
            out[name.value.decode()] = dict(launches=la.value, ms=ms.value, alg_bytes=by.value)
        return out

    def launch_count(self):
        return int(self.lib.vdn_launch_count(self.h))

    def comm_bytes(self):
        return int(self.lib.vdn_comm_bytes(self.h))
