#!/usr/bin/env python3
"""
tests/golden/make_golden.py -- regenerates tests/golden/*.npz.  Needs /root/reference (run `make -C oracle ref` first).

Each fixture holds, for one small seeded problem, the path-boundary inputs and the output of EVERY stage of the hot path
as computed by the reference's own per-box routines (oracle/_ref/libref.so = oracle/f2c.py's transpile of
src/{slope,velpred,mkflux,update,multifab_physbc,mkforce,make_at_halftime,macproject}.f90), each stage fed the inputs
listed in tests/test_golden.py.  `phi` (the MAC solve, FBoxLib F_MG, absent from the reference tree) and the projected
`umac` that the downstream stages consume come from the oracle's multigrid converged to 1e-13; they are INPUTS of the
fixture, not reference outputs.  tests/test_golden.py checks the CPU oracle (always) and the CUDA path (-m gpu) against
these files, so the reference pin travels to machines where /root/reference does not exist.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O, ref as R      # noqa: E402

W, NS, IN, OUT, PER, SYM = O.SLIP_WALL, O.NO_SLIP_WALL, O.INLET, O.OUTLET, O.PERIODIC, O.SYMMETRY

# name -> (kind, n, dim, max_grid_size, phys_bc, seed, param overrides)
CASES = {
    "g3d_mixed_2box": ("random", [8, 4, 4], 3, 4, [[IN, OUT], [W, NS], [OUT, IN]], 11, {}),
    "g3d_rt_1box": ("rt", [6, 6, 8], 3, 8, None, 0, {}),
    "g3d_per_so2_minion": ("random", [6, 6, 6], 3, 8, [[PER, PER], [W, W], [NS, NS]], 12,
                           dict(slope_order=2, use_minion=True, boussinesq=1)),
    "g2d_walls_4box": ("random", [16, 16], 2, 8, [[NS, NS], [NS, W]], 13, {}),
    "g2d_inout_so0": ("random", [12, 8], 2, 8, [[IN, OUT], [PER, PER]], 14, dict(slope_order=0)),
}


def build_case(name):
    kind, n, dim, mgs, bc, seed, over = CASES[name]
    rng = np.random.default_rng(seed)
    bcval = np.zeros((5, 3, 2))
    bcval[0:3] = rng.uniform(-0.5, 0.5, size=(3, 3, 2))
    bcval[3] = rng.uniform(1.0, 2.0, size=(3, 2))
    bcval[4] = rng.uniform(0.0, 1.0, size=(3, 2))
    P = O.Params(dim=dim, nscal=2, bcval=bcval, **over)
    if kind == "rt":
        geom, P, st, dt = O.rt_state(n, dim=dim, max_grid_size=mgs, params=P)
    else:
        geom, P, st, dt = O.random_state(n, dim=dim, max_grid_size=mgs, phys_bc=bc, seed=seed, params=P)
    return geom, P, st, dt


def main():
    if not R.available():
        raise SystemExit("oracle/_ref/libref.so missing: run `make -C oracle ref` (needs /root/reference)")
    for name in CASES:
        geom, P, st, dt = build_case(name)
        o = O.stagewise(geom, P, st, dt, mac_rel_eps=1e-13)
        r = R.stagewise_from(geom, P, st, dt, o)
        # the fixture is only valid if every stage input taken from the oracle equals the reference's output for it
        for k in R.PIN_KEYS:
            md, mb, nd = R.max_diff(o[k], r[k])
            assert nd == 0 or (k == "umac" and md <= 4e-16 * max(mb, 1.0)), (name, k, md, nd)
        meta = dict(name=name, dim=geom.dim, n_cell=geom.n_cell[:geom.dim], max_grid_size=CASES[name][3],
                    phys_bc=geom.phys_bc.tolist(), dt=dt, nscal=P.nscal, slope_order=P.slope_order, use_minion=P.use_minion,
                    boussinesq=P.boussinesq, bcval=P.bcval.tolist(), boxes=[[list(map(int, b[0])), list(map(int, b[1]))] for b in geom.boxes],
                    reference_routines={k: R.where(k) for k in ("velpred_3d", "velpred_2d", "mkflux_3d", "mkflux_2d", "update_3d",
                                                                "update_2d", "physbc_3d", "physbc_2d", "mkvelforce_3d",
                                                                "mkscalforce_3d", "make_at_halftime_3d", "divumac_3d",
                                                                "mk_mac_coeffs_3d", "mkumac_3d", "slopex_2d", "slopez_3d")})
        arrs = {}

        def put(prefix, mf):
            flat = R.flatten(mf)
            for i, a in enumerate(flat):
                arrs["%s/%d" % (prefix, i)] = a
        for k in ("uold", "sold", "gp", "ext_vel_force", "ext_scal_force"):
            put("in/" + k, st[k])
        put("in/phi", o["phi"])
        put("in/umac", o["umac"])                       # oracle-projected MAC velocity: input of mkflux/update
        for k in R.PIN_KEYS + ["rh", "beta"]:
            put("ref/" + k, r[k])
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **arrs)
        print("%-22s %6.1f KB  %d arrays" % (name, os.path.getsize(path) / 1024.0, len(arrs)))


def main_estdt():
    """tests/golden/estdt.json: the reference's own estdt_2d / estdt_3d (estdt.f90:89-181) inside the driver logic of estdt.f90:15-87, on the stored
    inputs of every fixture (and on a copy scaled below the reference's eps: the min(dx) fallback), as hexadecimal floats"""
    if not R.available():
        raise SystemExit("oracle/_ref/libref.so missing: run `make -C oracle ref` (needs /root/reference)")
    out = {"reference_routines": {k: R.where(k) for k in ("estdt_2d", "estdt_3d")}, "cases": {}}
    for name in CASES:
        geom, P, st, dt = build_case(name)
        rows = []
        for scale in (1.0, 1e-12):
            u = [a * scale for a in st["uold"]]; gp = [a * scale for a in st["gp"]]; f = [a * scale for a in st["ext_vel_force"]]
            for dtold in (-1.0, 1e-4):
                rows.append(dict(scale=scale, dtold=dtold, dt=float(R.estdt(geom, u, 3, st["sold"], 3, gp, 1, f, 1, dtold=dtold)).hex()))
        out["cases"][name] = rows
    with open(os.path.join(HERE, "estdt.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print("estdt.json: %d cases" % len(out["cases"]))


if __name__ == "__main__":
    if "--estdt" in sys.argv:
        main_estdt()
    else:
        main()
