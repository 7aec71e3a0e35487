"""
CPU test of the single-phase ghost-layer exchange of the multigrid level arrays (vdn_comm.cu: comm_halo_deep).  The message plan
comes from the library's own host-only planner (vdn_halo_plan, the function comm_halo_deep executes); this test plays every rank of a
process grid, moves the messages through per-pair FIFO queues (how NCCL matches the sends and receives of two ranks inside a
group) and checks every ghost cell a rank needs -- faces, edges, corners, and index n of the directions that are not split (where
the level arrays keep the high-boundary / periodic-seam face coefficient) -- against the global array.
Replaces multifab_fill_boundary (FBoxLib; call sites inside F_MG) between ranks for the fused smoother.
"""
import ctypes as C
from collections import defaultdict, deque

import numpy as np
import pytest

import varden_b200 as V

PAD = 4


def plan(lib, pgrid, pcoord, periodic, n, ng, dmask):
    I3 = C.c_int * 3
    nranks = pgrid[0] * pgrid[1] * pgrid[2]
    c2r = (C.c_int * nranks)(*range(nranks))          # rank = x + px*(y + py*z)
    ns, nr = C.c_int(0), C.c_int(0)
    sp, rp = (C.c_int * 26)(), (C.c_int * 26)()
    slo, sn, rlo, rn = (C.c_int * 78)(), (C.c_int * 78)(), (C.c_int * 78)(), (C.c_int * 78)()
    rc = lib.vdn_halo_plan(3, I3(*pgrid), I3(*pcoord), I3(*periodic), c2r, I3(*n), ng, dmask,
                           C.byref(ns), sp, slo, sn, C.byref(nr), rp, rlo, rn)
    assert rc == 0
    sends = [(sp[q], tuple(slo[3 * q:3 * q + 3]), tuple(sn[3 * q:3 * q + 3])) for q in range(ns.value)]
    recvs = [(rp[q], tuple(rlo[3 * q:3 * q + 3]), tuple(rn[3 * q:3 * q + 3])) for q in range(nr.value)]
    return sends, recvs


@pytest.mark.parametrize("pgrid,periodic", [
    ((2, 2, 2), (1, 1, 1)),      # 8 GPUs, all periodic: every diagonal neighbour is the same rank for several offsets
    ((2, 2, 2), (1, 1, 0)),      # bench.py --gpus 8: periodic x,y, walls in z
    ((2, 2, 2), (0, 1, 0)),      # the rand3d multi-GPU case at 8 ranks: inflow / outflow x, periodic y, walls z
    ((1, 2, 2), (1, 1, 0)),      # bench.py --gpus 4
    ((1, 1, 2), (1, 1, 0)),      # bench.py --gpus 2
    ((1, 1, 2), (1, 1, 1)),      # two ranks along a periodic direction: lo and hi neighbour are the same rank
    ((2, 1, 1), (1, 0, 1)),
    ((2, 2, 1), (0, 1, 1)),
    ((3, 2, 1), (1, 0, 0)),
    ((4, 1, 2), (0, 1, 1)),
])
@pytest.mark.parametrize("ng", [1, 3, 4])
def test_single_phase_plan_fills_every_ghost_cell(pgrid, periodic, ng):
    lib = V.load_library()
    n = (6, 4, 4)
    dmask = sum(1 << d for d in range(3) if pgrid[d] > 1)
    N = [n[d] * pgrid[d] for d in range(3)]

    def gval(i, j, k):                      # global level array; index N along a periodic direction is the seam face = index 0
        idx = [i, j, k]
        for d in range(3):
            if periodic[d]:
                idx[d] %= N[d]
        return float(idx[0] + 100 * idx[1] + 10000 * idx[2])

    ranks = [(x, y, z) for z in range(pgrid[2]) for y in range(pgrid[1]) for x in range(pgrid[0])]
    loc = {}
    for r, pc in enumerate(ranks):
        a = np.full(tuple(n[d] + 2 * PAD for d in (2, 1, 0)), np.nan)
        for k in range(n[2] + 1):
            for j in range(n[1] + 1):
                for i in range(n[0] + 1):
                    # a rank knows its own cells, and index n (the high-face coefficient) along the directions where no neighbour rank holds it:
                    # not split, or split with this rank at the physical high end of the domain
                    if all(idx < n[d] or (idx == n[d] and (pgrid[d] == 1 or (pc[d] == pgrid[d] - 1 and not periodic[d]))) for d, idx in enumerate((i, j, k))):
                        a[k + PAD, j + PAD, i + PAD] = gval(pc[0] * n[0] + i, pc[1] * n[1] + j, pc[2] * n[2] + k)
        loc[r] = a
    plans = {r: plan(lib, pgrid, ranks[r], periodic, n, ng, dmask) for r in range(len(ranks))}
    sl = lambda lo, nn: tuple(slice(lo[d] + PAD, lo[d] + PAD + nn[d]) for d in (2, 1, 0))
    fifo = defaultdict(deque)
    for r, (sends, _) in plans.items():
        for peer, lo, nn in sends:
            assert peer != r or pgrid[0] * pgrid[1] * pgrid[2] == 1
            fifo[(r, peer)].append(loc[r][sl(lo, nn)].copy())
    for r, (_, recvs) in plans.items():
        for peer, lo, nn in recvs:
            buf = fifo[(peer, r)].popleft()
            assert buf.shape == loc[r][sl(lo, nn)].shape
            loc[r][sl(lo, nn)] = buf
    assert all(len(q) == 0 for q in fifo.values()), "unmatched messages"
    checked = 0
    for r, pc in enumerate(ranks):
        rng = []
        for d in range(3):
            if pgrid[d] > 1:
                lo = -ng if (pc[d] > 0 or periodic[d]) else 0
                hi = n[d] + ng if (pc[d] < pgrid[d] - 1 or periodic[d]) else n[d] + 1      # physical high end: index n (the boundary face) included
            else:
                lo, hi = 0, n[d] + 1
            rng.append(range(lo, hi))
        for k in rng[2]:
            for j in rng[1]:
                for i in rng[0]:
                    want = gval(pc[0] * n[0] + i, pc[1] * n[1] + j, pc[2] * n[2] + k)
                    assert loc[r][k + PAD, j + PAD, i + PAD] == want, (r, i, j, k)
                    checked += 1
    assert checked > 0


def plan_ex(lib, pgrid, pcoord, periodic, n, ng, dmask, nodal, carry_n):
    I3 = C.c_int * 3
    nranks = pgrid[0] * pgrid[1] * pgrid[2]
    c2r = (C.c_int * nranks)(*range(nranks))
    ns, nr = C.c_int(0), C.c_int(0)
    sp, rp = (C.c_int * 26)(), (C.c_int * 26)()
    slo, sn, rlo, rn, rsh = ((C.c_int * 78)() for _ in range(5))
    rc = lib.vdn_halo_plan_ex(3, I3(*pgrid), I3(*pcoord), I3(*periodic), c2r, I3(*n), ng, dmask, nodal, carry_n,
                              C.byref(ns), sp, slo, sn, C.byref(nr), rp, rlo, rn, rsh)
    assert rc == 0
    sends = [(sp[q], tuple(slo[3 * q:3 * q + 3]), tuple(sn[3 * q:3 * q + 3])) for q in range(ns.value)]
    recvs = [(rp[q], tuple(rlo[3 * q:3 * q + 3]), tuple(rn[3 * q:3 * q + 3]), tuple(rsh[3 * q:3 * q + 3])) for q in range(nr.value)]
    return sends, recvs


@pytest.mark.parametrize("pgrid,periodic", [
    ((2, 2, 2), (1, 1, 0)),      # bench.py --gpus 8
    ((2, 2, 2), (0, 1, 0)),      # rand3d at 8 ranks
    ((1, 2, 2), (1, 1, 0)),      # bench.py --gpus 4
    ((1, 1, 2), (1, 1, 1)),      # lo and hi neighbour are the same rank
    ((2, 2, 2), (1, 1, 1)),
    ((3, 1, 2), (0, 1, 1)),
])
@pytest.mark.parametrize("nodal,ng", [(-1, 3), (-1, 1), (0, 1), (1, 1), (2, 1)])
@pytest.mark.parametrize("transport", ["nccl_fifo", "peer_pull"])
def test_field_fill_boundary_plan(pgrid, periodic, nodal, ng, transport):
    """multifab_fill_boundary of a cell- or face-centred FIELD over a process grid, as vdn_stream.cu:st_fill_boundary does it: one exchange over
    the split directions (both transports: messages matched first-in first-out per pair of ranks as NCCL does, or every rank reading the
    peers' arrays with the plan's index shift as k_halo_pull does), then the periodic directions a rank owns alone wrap in order over the
    ghosted range.  Every ghost cell / face that has a periodic or rank neighbour must equal the global array."""
    lib = V.load_library()
    n = (6, 4, 4)
    G = 3                                                       # storage ghost width
    dmask = sum(1 << d for d in range(3) if pgrid[d] > 1)
    N = [n[d] * pgrid[d] for d in range(3)]
    nod = [1 if d == nodal else 0 for d in range(3)]

    def gval(i, j, k):
        idx = [i, j, k]
        for d in range(3):
            if periodic[d]:
                idx[d] %= N[d]                                  # a periodic face N is face 0
        return float(idx[0] + 100 * idx[1] + 10000 * idx[2])

    ranks = [(x, y, z) for z in range(pgrid[2]) for y in range(pgrid[1]) for x in range(pgrid[0])]
    loc = {}
    for r, pc in enumerate(ranks):
        a = np.full(tuple(n[d] + nod[d] + 2 * G for d in (2, 1, 0)), np.nan)
        for k in range(n[2] + nod[2]):
            for j in range(n[1] + nod[1]):
                for i in range(n[0] + nod[0]):
                    a[k + G, j + G, i + G] = gval(pc[0] * n[0] + i, pc[1] * n[1] + j, pc[2] * n[2] + k)
        loc[r] = a
    plans = {r: plan_ex(lib, pgrid, ranks[r], periodic, n, ng, dmask, nodal, 0) for r in range(len(ranks))}
    sl = lambda lo, nn, sh=(0, 0, 0): tuple(slice(lo[d] + sh[d] + G, lo[d] + sh[d] + G + nn[d]) for d in (2, 1, 0))
    if transport == "nccl_fifo":
        fifo = defaultdict(deque)
        for r, (sends, _) in plans.items():
            for peer, lo, nn in sends:
                fifo[(r, peer)].append(loc[r][sl(lo, nn)].copy())
        for r, (_, recvs) in plans.items():
            for peer, lo, nn, sh in recvs:
                loc[r][sl(lo, nn)] = fifo[(peer, r)].popleft()
        assert all(len(q) == 0 for q in fifo.values())
    else:
        snap = {r: a.copy() for r, a in loc.items()}            # what the peers had published when the exchange started
        for r, (_, recvs) in plans.items():
            for peer, lo, nn, sh in recvs:
                src = snap[peer][sl(lo, nn, sh)]
                assert not np.isnan(src).any(), "a pull read cells its peer does not own"
                loc[r][sl(lo, nn)] = src
    # local periodic wraps, x then y then z (st_fill_boundary)
    for r, pc in enumerate(ranks):
        a = loc[r]
        for d in range(3):
            if not (periodic[d] and pgrid[d] == 1):
                continue
            rngs = []
            for t in range(3):
                filled = t != d and (pgrid[t] > 1 or (periodic[t] and pgrid[t] == 1 and t < d))
                rngs.append(range(-ng, n[t] + nod[t] + ng) if filled else range(0, n[t] + nod[t]))
            for g in range(1, ng + 1):
                for t2 in rngs[(d + 2) % 3]:
                    for t1 in rngs[(d + 1) % 3]:
                        def at(v):
                            idx = [0, 0, 0]; idx[d] = v; idx[(d + 1) % 3] = t1; idx[(d + 2) % 3] = t2
                            return (idx[2] + G, idx[1] + G, idx[0] + G)
                        if nod[d]:
                            a[at(-g)] = a[at(n[d] - g)]; a[at(n[d] + g)] = a[at(g)]
                        else:
                            a[at(-g)] = a[at(n[d] - g)]; a[at(n[d] - 1 + g)] = a[at(g - 1)]
    checked = 0
    for r, pc in enumerate(ranks):
        rng = []
        for d in range(3):
            lo = -ng if (pc[d] > 0 or periodic[d]) else 0
            hi = n[d] + nod[d] + ng if (pc[d] < pgrid[d] - 1 or periodic[d]) else n[d] + nod[d]
            rng.append(range(lo, hi))
        for k in rng[2]:
            for j in rng[1]:
                for i in rng[0]:
                    assert loc[r][k + G, j + G, i + G] == gval(pc[0] * n[0] + i, pc[1] * n[1] + j, pc[2] * n[2] + k), (r, i, j, k)
                    checked += 1
    assert checked > 0
