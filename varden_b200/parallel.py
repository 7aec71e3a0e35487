"""
varden_b200/parallel.py -- host-side partitioning of one level's boxes over ranks (one rank per GPU) and the
communicator bootstrap.  Replaces FBoxLib's layout_build_ba + parallel (MPI) for this path: boxes are dealt out in
contiguous rectangular blocks so every rank owns a rectangular region of equal size; the NCCL unique id is created by
rank 0 inside libvdn.so and broadcast through torch.distributed (plumbing only).
"""
import ctypes as C

import numpy as np


def process_grid(world, dim, nb):
    """factorise `world` over the box grid nb[3], filling the LAST direction first (z, then y, then x)"""
    pg = [1, 1, 1]
    w = world
    order = list(range(dim - 1, -1, -1))
    while w > 1:
        placed = False
        for d in order:
            if w % 2 == 0 and nb[d] % (pg[d] * 2) == 0:
                # keep the grid as cubic as possible: prefer the direction with the smallest factor so far
                best = min((dd for dd in order if nb[dd] % (pg[dd] * 2) == 0), key=lambda dd: (pg[dd], -dd))
                pg[best] *= 2
                w //= 2
                placed = True
                break
        if not placed:
            raise ValueError("cannot factorise %d ranks over a %s box grid" % (world, nb))
    return pg


def partition(geom, world):
    """-> (box ids per rank, region_lo[world][3], region_hi[world][3], process grid)"""
    dim = geom.dim
    los = [sorted(set(b[0][d] for b in geom.boxes)) for d in range(3)]
    nb = [len(los[d]) for d in range(3)]
    if nb[0] * nb[1] * nb[2] != geom.nboxes:
        raise ValueError("boxes are not a tensor-product chop of the domain")
    pg = process_grid(world, dim, nb)
    per = [nb[d] // pg[d] for d in range(3)]
    owner = {}
    for ib, (lo, hi) in enumerate(geom.boxes):
        bc = [los[d].index(lo[d]) for d in range(3)]
        pc = [bc[d] // per[d] for d in range(3)]
        r = pc[0] + pg[0] * (pc[1] + pg[1] * pc[2])
        owner.setdefault(r, []).append(ib)
    ids = [owner.get(r, []) for r in range(world)]
    rlo = np.zeros((world, 3), dtype=np.int32)
    rhi = np.zeros((world, 3), dtype=np.int32)
    for r in range(world):
        if not ids[r]:
            raise ValueError("rank %d received no boxes" % r)
        rlo[r] = np.min([geom.boxes[i][0] for i in ids[r]], axis=0)
        rhi[r] = np.max([geom.boxes[i][1] for i in ids[r]], axis=0)
    return ids, rlo, rhi, pg


def comm_plan(geom, rank, world, rlo, rhi):
    """neighbour ranks [3][2] (-1 = physical boundary), process grid and this rank's coordinates (host-only C call)"""
    from . import load_library
    lib = load_library()
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    dlo = np.ascontiguousarray(geom.dlo, dtype=np.int32)
    dhi = np.ascontiguousarray(geom.dhi, dtype=np.int32)
    pbc = np.ascontiguousarray(geom.phys_bc, dtype=np.int32)
    nbr = np.zeros(6, dtype=np.int32)
    pg = np.zeros(3, dtype=np.int32)
    pc = np.zeros(3, dtype=np.int32)
    rlo = np.ascontiguousarray(rlo, dtype=np.int32)
    rhi = np.ascontiguousarray(rhi, dtype=np.int32)
    rc = lib.vdn_comm_plan(geom.dim, rank, world, ip(rlo), ip(rhi), ip(dlo), ip(dhi), ip(pbc), ip(nbr), ip(pg), ip(pc))
    if rc != 0:
        raise ValueError("vdn_comm_plan failed (%d): regions are not a tensor-product decomposition" % rc)
    return nbr.reshape(3, 2), pg, pc


def init_comm(ctx, rank, world, rlo, rhi):
    """create the NCCL communicator of a context; the 128-byte unique id travels through torch.distributed"""
    import torch.distributed as dist
    from . import load_library
    lib = load_library()
    buf = (C.c_ubyte * 128)()
    if rank == 0:
        if lib.vdn_nccl_unique_id(buf) != 0:
            raise RuntimeError("ncclGetUniqueId failed")
    obj = [bytes(buf)]
    dist.broadcast_object_list(obj, src=0)
    idb = (C.c_ubyte * 128).from_buffer_copy(obj[0])
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    rlo = np.ascontiguousarray(rlo, dtype=np.int32)
    rhi = np.ascontiguousarray(rhi, dtype=np.int32)
    rc = lib.vdn_ctx_set_comm(ctx.h, rank, world, ip(rlo), ip(rhi), idb)
    if rc != 0:
        raise RuntimeError(lib.vdn_last_error(ctx.h).decode())
