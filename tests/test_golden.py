"""
Golden-vector tests: tests/golden/*.npz hold stage-by-stage outputs of the reference's OWN per-box routines (transpiled
from /root/reference by oracle/f2c.py; generator: tests/golden/make_golden.py).  Checked here, on the stored inputs:
  * CPU (always): the hand-written oracle reproduces every stage BIT FOR BIT (mkumac's box-boundary faces, which the
    reference takes from F_MG's fine_flx, within 4 ulp);
  * GPU (-m gpu): the CUDA path through the C ABI, same stages, bar 1e-12 relative L-inf (north_star) -- observed 0.
"""
import glob
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))
TOL_EDGE = 1e-12


class Gold:
    def __init__(self, path):
        z = np.load(path)
        self.meta = json.loads(bytes(z["meta"]).decode())
        m = self.meta
        self.geom = O.Geom(m["dim"], m["n_cell"], m["phys_bc"], max_grid_size=m["max_grid_size"])
        assert [[list(b[0]), list(b[1])] for b in self.geom.boxes] == m["boxes"]
        self.P = O.Params(dim=m["dim"], nscal=m["nscal"], slope_order=m["slope_order"], use_minion=m["use_minion"],
                          boussinesq=m["boussinesq"], bcval=m["bcval"])
        self.dt = m["dt"]
        self.z = z
        self.nb = self.geom.nboxes

    def mf(self, key):
        return [np.asfortranarray(self.z["%s/%d" % (key, i)]) for i in range(self.nb)]

    def mfd(self, key):
        return [[np.asfortranarray(self.z["%s/%d" % (key, d * self.nb + i)]) for i in range(self.nb)] for d in range(self.geom.dim)]

    def state(self):
        return {k: self.mf("in/" + k) for k in ("uold", "sold", "gp", "ext_vel_force", "ext_scal_force")}


def test_fixtures_present():
    assert len(GOLD) >= 5


def _exact(name, got, ref):
    from oracle.ref import max_diff
    md, mb, nd = max_diff(got, ref)
    assert nd == 0, "%s: %d entries differ from the reference, max |diff| %.3e (max |ref| %.3e)" % (name, nd, md, mb)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_matches_reference_golden(path):
    g = Gold(path)
    geom, P, dt, st = g.geom, g.P, g.dt, g.state()
    dim, nscal = geom.dim, P.nscal
    A = O.mf_alloc
    lapu, laps, divu, mac_rhs = A(geom, 0, dim), A(geom, 0, nscal), A(geom, 1, 1), A(geom, 1, 1)

    vf1 = A(geom, 1, dim)
    O.mkvelforce(geom, P, vf1, st["ext_vel_force"], st["gp"], st["sold"], 3, nscal, lapu, 1.0)
    _exact("vel_force_1", vf1, g.mf("ref/vel_force_1"))

    up = [A(geom, 1, 1, d, val=1.0e20) for d in range(dim)]
    O.velpred(geom, P, st["uold"], up, g.mf("ref/vel_force_1"), dt)
    _exact("umac_pred", up, g.mfd("ref/umac_pred"))

    # macproject glue on the stored phi: interior faces exact, box-boundary faces within a few ulp (fine_flx scaling)
    um = O.project_with_phi(geom, P, g.mfd("ref/umac_pred"), st["sold"], nscal, g.mf("in/phi"))
    from oracle.ref import max_diff
    md, mb, nd = max_diff(um, g.mfd("ref/umac"))
    assert md <= 4e-16 * max(mb, 1.0), ("umac", md, nd)

    umac = g.mfd("in/umac")
    sf1 = A(geom, 1, nscal)
    O.mkscalforce(geom, P, sf1, st["ext_scal_force"], laps, 1.0)
    _exact("scal_force_1", sf1, g.mf("ref/scal_force_1"))
    sedge = [A(geom, 0, nscal, d) for d in range(dim)]
    sflux = [A(geom, 0, nscal, d) for d in range(dim)]
    ics = [1] + [0] * (nscal - 1)
    O.mkflux(geom, P, st["sold"], nscal, sedge, sflux, umac, sf1, divu, dt, False, ics)
    _exact("sedge", sedge, g.mfd("ref/sedge"))
    _exact("sflux", sflux, g.mfd("ref/sflux"))
    sf2 = A(geom, 1, nscal)
    O.mkscalforce(geom, P, sf2, st["ext_scal_force"], laps, 0.0)
    _exact("scal_force_2", sf2, g.mf("ref/scal_force_2"))
    snew = A(geom, 3, nscal)
    O.update(geom, P, st["sold"], nscal, umac, g.mfd("ref/sedge"), g.mfd("ref/sflux"), sf2, snew, dt, False, ics)
    _exact("snew", snew, g.mf("ref/snew"))
    rhoh = A(geom, 1, 1)
    O.make_at_halftime(geom, P, rhoh, st["sold"], g.mf("ref/snew"))
    _exact("rhohalf", rhoh, g.mf("ref/rhohalf"))
    uedge = [A(geom, 0, dim, d) for d in range(dim)]
    uflux = [A(geom, 0, dim, d) for d in range(dim)]
    O.mkflux(geom, P, st["uold"], dim, uedge, uflux, umac, g.mf("ref/vel_force_1"), mac_rhs, dt, True, [0] * dim)
    _exact("uedge", uedge, g.mfd("ref/uedge"))
    vf2 = A(geom, 1, dim)
    O.mkvelforce(geom, P, vf2, st["ext_vel_force"], st["gp"], g.mf("ref/rhohalf"), 1, 1, lapu, 0.0)
    _exact("vel_force_2", vf2, g.mf("ref/vel_force_2"))
    unew = A(geom, 3, dim)
    O.update(geom, P, st["uold"], dim, umac, g.mfd("ref/uedge"), uflux, vf2, unew, dt, True, [0] * dim)
    _exact("unew", unew, g.mf("ref/unew"))


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_cuda_matches_reference_golden(path):
    from util import make_ctx, upload_state, relerr, download_like
    g = Gold(path)
    geom, P, dt, st = g.geom, g.P, g.dt, g.state()
    dim, nscal = geom.dim, P.nscal
    ctx = make_ctx(geom, P)
    upload_state(ctx, geom, P, st)
    errs = {}

    def chk(key, field, ref, ng, nc, face_dir=-1):
        got = download_like(ctx, geom, field, ref, ng, nc)
        errs[key] = relerr(geom, got, ref, ng, face_dir)

    ctx.mkvelforce("SOLD", 1.0)
    chk("vel_force_1", "VEL_FORCE", g.mf("ref/vel_force_1"), 1, dim)
    ctx.velpred(dt)
    up = g.mfd("ref/umac_pred")
    for d in range(dim):
        chk("umac_pred%d" % d, "UMAC_" + "XYZ"[d], up[d], 1, 1, d)
    umac = g.mfd("in/umac")
    for d in range(dim):
        ctx.upload_mf("UMAC_" + "XYZ"[d], umac[d], 1, 1)
    ctx.mkscalforce(1.0)
    ctx.mkflux(False, dt)
    se, sfl = g.mfd("ref/sedge"), g.mfd("ref/sflux")
    for d in range(dim):
        chk("sedge%d" % d, "SEDGE_" + "XYZ"[d], se[d], 0, nscal, d)
        chk("sflux%d" % d, "SFLUX_" + "XYZ"[d], [a[..., :1].copy(order="F") for a in sfl[d]], 0, 1, d)
    ctx.mkscalforce(0.0)
    ctx.update(False, dt)
    chk("snew", "SNEW", g.mf("ref/snew"), 3, nscal)
    ctx.make_at_halftime()
    chk("rhohalf", "RHOHALF", g.mf("ref/rhohalf"), 1, 1)
    ctx.mkvelforce("SOLD", 1.0)
    ctx.mkflux(True, dt)
    ue = g.mfd("ref/uedge")
    for d in range(dim):
        chk("uedge%d" % d, "UEDGE_" + "XYZ"[d], ue[d], 0, dim, d)
    ctx.mkvelforce("RHOHALF", 0.0)
    chk("vel_force_2", "VEL_FORCE", g.mf("ref/vel_force_2"), 1, dim)
    ctx.update(True, dt)
    chk("unew", "UNEW", g.mf("ref/unew"), 3, dim)
    ctx.close()
    print(os.path.basename(path), {k: "%.1e" % v for k, v in errs.items()})
    bad = {k: v for k, v in errs.items() if not (v <= TOL_EDGE)}
    assert not bad, bad


def _estdt_gold():
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "estdt.json")
    return json.load(open(p))["cases"]


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_estdt_matches_reference_golden(path):
    """estdt (estdt.f90:15-181): tests/golden/estdt.json holds the reference routines' dt on the fixture inputs"""
    g = Gold(path)
    st = g.state()
    for row in _estdt_gold()[g.meta["name"]]:
        sc = row["scale"]
        u = [a * sc for a in st["uold"]]; gp = [a * sc for a in st["gp"]]; f = [a * sc for a in st["ext_vel_force"]]
        got = O.estdt(g.geom, u, 3, st["sold"], 3, gp, 1, f, 1, dtold=row["dtold"])
        assert got == float.fromhex(row["dt"]), (g.meta["name"], row, got)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_cuda_estdt_matches_reference_golden(path):
    from util import make_ctx, upload_state
    g = Gold(path)
    st = g.state()
    ctx = make_ctx(g.geom, g.P)
    for row in _estdt_gold()[g.meta["name"]]:
        sc = row["scale"]
        s2 = dict(st)
        for k in ("uold", "gp", "ext_vel_force"):
            s2[k] = [np.asfortranarray(a * sc) for a in st[k]]
        upload_state(ctx, g.geom, g.P, s2)
        got = ctx.estdt(dtold=row["dtold"])
        assert got == float.fromhex(row["dt"]), (g.meta["name"], row, got)
    ctx.close()
