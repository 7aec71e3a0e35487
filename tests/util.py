"""Shared helpers of the parity tests: build a device context from an oracle Geom/Params and compare multifabs."""
import numpy as np

from oracle import oracle as O
import varden_b200 as V


def make_ctx(geom, P, **kw):
    prm = V.default_params(nscal=P.nscal, slope_order=P.slope_order, use_minion=P.use_minion, boussinesq=P.boussinesq,
                           visc_coef=P.visc_coef, diff_coef=P.diff_coef, bc_val=P.bcval, **kw)
    return V.Context(geom.dim, geom.boxes, geom.dlo, geom.dhi, geom.phys_bc, geom.dx, params=prm)


def upload_state(ctx, geom, P, st):
    dim = geom.dim
    ctx.upload_mf("UOLD", st["uold"], 3, dim)
    ctx.upload_mf("SOLD", st["sold"], 3, P.nscal)
    ctx.upload_mf("GP", st["gp"], 1, dim)
    ctx.upload_mf("EXT_VEL_FORCE", st["ext_vel_force"], 1, dim)
    ctx.upload_mf("EXT_SCAL_FORCE", st["ext_scal_force"], 1, P.nscal)


def relerr(geom, got, ref, ng, face_dir=-1, comps=None, full=None):
    """relative L-inf error over the valid region of every box (and over the whole arrays for single-box layouts)"""
    if full is None:
        full = geom.nboxes == 1
    num, den = 0.0, 0.0
    for ib in range(geom.nboxes):
        a, b = got[ib], ref[ib]
        if not full:
            a, b = O.valid(geom, a, ib, ng, face_dir), O.valid(geom, b, ib, ng, face_dir)
        if comps is not None:
            a, b = a[..., comps], b[..., comps]
        if not np.all(np.isfinite(b)):                   # a non-finite reference means the test itself is broken
            return np.inf
        m = np.isfinite(b) & (np.abs(b) < 1e19)          # skip the 1.d20 poison in umac ghost faces (compared separately)
        if not np.array_equal(np.abs(b) >= 1e19, np.abs(a) >= 1e19):
            return np.inf
        if m.any():
            d = np.abs(a[m] - b[m])
            if not np.all(np.isfinite(d)):          # NaN/inf anywhere (e.g. cells a download did not write) is a failure
                return np.inf
            num = max(num, float(d.max()))
            den = max(den, float(np.abs(b[m]).max()))
    return num / den if den > 0 else num


def download_like(ctx, geom, field, ref, ng, ncomp):
    out = [np.full_like(a, np.nan) for a in ref]      # NaN prefill: anything the download does not write stays visible
    ctx.download_mf(field, out, ng, ncomp)
    return out
