"""
The drop-in boundary driven by a COMPILED C stand-in for the Fortran driver (tests/mock_driver.c), not by ctypes: it owns
Fortran-layout host "multifabs" and replays advance_timestep.f90:95-124 through the same-named procedures of
fortran/vdn_modules.f90 (velpred, macproject, mkflux, update: inputs copied in, stage, outputs copied out).  Its outputs are
checked against the golden fixtures of the reference's own routines (tests/golden/, bar 1e-12; the projected MAC velocity
against the stored one to 10x the solver tolerance).
"""
import os
import struct
import subprocess
import tempfile

import numpy as np
import pytest

from test_golden import GOLD, Gold

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(ROOT, "varden_b200", "libvdn.so")
TOL_EDGE, TOL_MAC = 1e-12, 1e-9


@pytest.fixture(scope="module")
def exe():
    out = os.path.join(HERE, "emu", "mock_driver")
    src = os.path.join(HERE, "mock_driver.c")
    hdr = os.path.join(ROOT, "include", "vdn.h")
    if not os.path.exists(out) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(out):
        subprocess.check_call(["gcc", "-O1", "-Wall", "-Werror", "-std=c11", "-D_GNU_SOURCE", src, "-o", out, "-ldl"])
    return out


def write_deck(path, g, use_given_umac=1):
    geom, P = g.geom, g.P
    st = g.state()
    nb, dim = geom.nboxes, geom.dim
    with open(path, "wb") as f:
        f.write(struct.pack("6i", dim, nb, P.nscal, P.slope_order, int(P.use_minion), int(P.boussinesq)))
        lo = np.zeros((nb, 3), np.int32); hi = np.zeros((nb, 3), np.int32)
        for i, (l, h) in enumerate(geom.boxes):
            lo[i, :dim] = l[:dim]; hi[i, :dim] = h[:dim]
        f.write(lo.tobytes()); f.write(hi.tobytes())
        d3 = lambda v: np.array(list(v[:dim]) + [0] * (3 - dim), np.int32)
        f.write(d3(geom.dlo).tobytes()); f.write(d3(geom.dhi).tobytes())
        pbc = np.zeros((3, 2), np.int32); pbc[:dim] = np.asarray(geom.phys_bc)[:dim]
        f.write(pbc.tobytes())
        dx = np.ones(3); dx[:dim] = geom.dx[:dim]
        f.write(dx.tobytes()); f.write(struct.pack("d", g.dt))
        f.write(np.ascontiguousarray(np.asarray(P.bcval, dtype=np.float64).reshape(5, 3, 2)).tobytes())
        f.write(struct.pack("i", use_given_umac))
        for k in ("uold", "sold", "gp", "ext_vel_force", "ext_scal_force"):
            for a in st[k]:
                f.write(np.asfortranarray(a).tobytes(order="F"))
        for d in range(dim):
            for a in g.mfd("in/umac")[d]:
                f.write(np.asfortranarray(a).tobytes(order="F"))


def test_mock_driver_builds_and_refuses_without_gpu(exe):
    """the driver compiles against include/vdn.h; on a machine without a CUDA device the library must refuse loudly (no CPU fallback)"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the parity test")
    g = Gold(GOLD[0])
    with tempfile.TemporaryDirectory() as td:
        deck, out = os.path.join(td, "deck.bin"), os.path.join(td, "out.bin")
        write_deck(deck, g)
        r = subprocess.run([exe, LIB, deck, out], capture_output=True, text=True, timeout=120)
        assert r.returncode != 0
        assert "CUDA" in r.stderr or "cuda" in r.stderr, r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_mock_driver_matches_reference_golden(exe, path):
    g = Gold(path)
    geom, P = g.geom, g.P
    dim, nscal, nb = geom.dim, P.nscal, geom.nboxes
    with tempfile.TemporaryDirectory() as td:
        deck, out = os.path.join(td, "deck.bin"), os.path.join(td, "out.bin")
        write_deck(deck, g)
        r = subprocess.run([exe, LIB, deck, out], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        raw = np.fromfile(out, dtype=np.float64, count=(os.path.getsize(out) - 12) // 8)
    pos = [0]

    def take(like):
        res = []
        for a in like:
            n = a.size
            res.append(raw[pos[0]:pos[0] + n].reshape(a.shape, order="F")); pos[0] += n
        return res

    def err(got, ref):
        num = max(float(np.abs(a - b)[np.abs(b) < 1e19].max()) for a, b in zip(got, ref))
        den = max(float(np.abs(b)[np.abs(b) < 1e19].max()) for b in ref)
        return num / den if den > 0 else num

    from oracle import oracle as O
    errs = {}
    errs["vel_force_1"] = err(take(g.mf("ref/vel_force_1")), g.mf("ref/vel_force_1"))
    up = g.mfd("ref/umac_pred")
    for d in range(dim):
        got = take(up[d])
        errs["umac_pred%d" % d] = max(np.abs(O.valid(geom, a, ib, 1, d) - O.valid(geom, b, ib, 1, d)).max() for ib, (a, b) in enumerate(zip(got, up[d])))
    um = g.mfd("ref/umac")
    scale = max(np.abs(O.valid(geom, b, ib, 1, d)).max() for d in range(dim) for ib, b in enumerate(um[d]))
    mac = {}
    for d in range(dim):
        got = take(um[d])
        mac["umac%d" % d] = max(np.abs(O.valid(geom, a, ib, 1, d) - O.valid(geom, b, ib, 1, d)).max() for ib, (a, b) in enumerate(zip(got, um[d]))) / scale
    se, sfl = g.mfd("ref/sedge"), g.mfd("ref/sflux")
    for d in range(dim):
        errs["sedge%d" % d] = err(take(se[d]), se[d])
        errs["sflux%d" % d] = err(take(sfl[d]), sfl[d])
    errs["snew"] = err([O.valid(geom, a, ib, 3) for ib, a in enumerate(take(g.mf("ref/snew")))], [O.valid(geom, a, ib, 3) for ib, a in enumerate(g.mf("ref/snew"))])
    rh_ref = g.mf("ref/rhohalf")
    rh_like = [np.zeros(a.shape[:3] + (dim,), order="F") for a in rh_ref]
    rh = take(rh_like)
    errs["rhohalf"] = err([O.valid(geom, a[..., :1], ib, 1) for ib, a in enumerate(rh)], [O.valid(geom, a[..., :1], ib, 1) for ib, a in enumerate(rh_ref)])
    ue = g.mfd("ref/uedge")
    for d in range(dim):
        errs["uedge%d" % d] = err(take(ue[d]), ue[d])
    errs["vel_force_2"] = err(take(g.mf("ref/vel_force_2")), g.mf("ref/vel_force_2"))
    errs["unew"] = err([O.valid(geom, a, ib, 3) for ib, a in enumerate(take(g.mf("ref/unew")))], [O.valid(geom, a, ib, 3) for ib, a in enumerate(g.mf("ref/unew"))])
    print(os.path.basename(path), r.stdout.strip(), {k: "%.1e" % v for k, v in errs.items()}, {k: "%.1e" % v for k, v in mac.items()})
    assert not {k: v for k, v in errs.items() if not (v <= TOL_EDGE)}, errs
    assert not {k: v for k, v in mac.items() if not (v <= TOL_MAC)}, mac
