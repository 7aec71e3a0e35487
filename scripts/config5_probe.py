#!/usr/bin/env python
"""BASELINE config 5 probe (512^3, density ratio 1000:1): the per-cycle residual history of the MAC solve (mg_verbose), to see where
the relative residual stalls -- the FP64 floor of |r|_inf / |rh|_inf grows with the coefficient ratio and with n^2.
  python scripts/config5_probe.py [n_global] [ratio] [max_cycles]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import varden_b200 as V
from varden_b200.problems import rt_problem, Geom, PERIODIC, NO_SLIP_WALL

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ratio = float(sys.argv[2]) if len(sys.argv) > 2 else 1000.0
maxc = int(sys.argv[3]) if len(sys.argv) > 3 else 45
mgs = min(n, 256)
phi = [float(n) / mgs] * 3
g = Geom(3, [n] * 3, [[PERIODIC, PERIODIC], [PERIODIC, PERIODIC], [NO_SLIP_WALL, NO_SLIP_WALL]], prob_hi=phi, max_grid_size=mgs)
geom, st, dt = rt_problem([n] * 3, dim=3, max_grid_size=mgs, ratio=ratio, prob_hi=phi, box_ids=list(range(g.nboxes)))
ctx = V.Context(3, geom.boxes, g.dlo, g.dhi, g.phys_bc, g.dx, params=V.default_params(mg_verbose=1, mg_max_cycles=maxc))
for fld, key, ng, nc in (("UOLD", "uold", 3, 3), ("SOLD", "sold", 3, 2), ("GP", "gp", 1, 3), ("EXT_VEL_FORCE", "ext_vel_force", 1, 3),
                         ("EXT_SCAL_FORCE", "ext_scal_force", 1, 2)):
    ctx.upload_mf(fld, st[key], ng, nc)
ctx.fill_and_physbc("UOLD", 0)
ctx.fill_and_physbc("SOLD", 3)
ctx.fill_boundary("GP")
try:
    print("config5 probe n=%d ratio=%g:" % (n, ratio), ctx.advance(dt), flush=True)
except V.VdnError as e:
    print("config5 probe n=%d ratio=%g: %s" % (n, ratio, e), flush=True)
ctx.close()
