"""Host-side multi-rank logic on CPU: partition of boxes over ranks and the neighbour plan, cross-checked between
two real processes over the gloo backend (world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from varden_b200.problems import Geom, PERIODIC, NO_SLIP_WALL, SLIP_WALL
from varden_b200 import parallel as PAR


def test_partition_blocks_are_rectangular_and_equal():
    geom = Geom(3, [64, 64, 64], [[PERIODIC, PERIODIC], [PERIODIC, PERIODIC], [NO_SLIP_WALL, NO_SLIP_WALL]], max_grid_size=32)
    for world, want in ((1, [1, 1, 1]), (2, [1, 1, 2]), (4, [1, 2, 2]), (8, [2, 2, 2])):
        ids, rlo, rhi, pg = PAR.partition(geom, world)
        assert pg == want
        assert sorted(i for r in ids for i in r) == list(range(geom.nboxes))
        sizes = {tuple(rhi[r] - rlo[r] + 1) for r in range(world)}
        assert len(sizes) == 1
        for r in range(world):
            cells = sum(np.prod(np.array(geom.boxes[i][1]) - np.array(geom.boxes[i][0]) + 1) for i in ids[r])
            assert cells == np.prod(rhi[r] - rlo[r] + 1)


def test_comm_plan_periodic_and_walls():
    geom = Geom(3, [64, 64, 64], [[PERIODIC, PERIODIC], [SLIP_WALL, SLIP_WALL], [NO_SLIP_WALL, NO_SLIP_WALL]], max_grid_size=32)
    ids, rlo, rhi, pg = PAR.partition(geom, 8)
    for r in range(8):
        nbr, g, pc = PAR.comm_plan(geom, r, 8, rlo, rhi)
        assert list(g) == [2, 2, 2]
        # periodic x with 2 ranks: both neighbours are the other rank in x
        assert nbr[0, 0] == nbr[0, 1] != r
        # walls: no neighbour beyond the domain
        assert (nbr[1, 0] == -1) == (pc[1] == 0) and (nbr[1, 1] == -1) == (pc[1] == 1)
        assert (nbr[2, 0] == -1) == (pc[2] == 0) and (nbr[2, 1] == -1) == (pc[2] == 1)
    # a single rank along a periodic direction is its own neighbour (index wrap on the device)
    ids, rlo, rhi, pg = PAR.partition(geom, 2)
    nbr, g, pc = PAR.comm_plan(geom, 0, 2, rlo, rhi)
    assert nbr[0, 0] == 0 and nbr[0, 1] == 0 and nbr[2, 1] == 1 and nbr[2, 0] == -1


def test_comm_plan_rejects_bad_decomposition():
    geom = Geom(3, [64, 64, 64], [[PERIODIC, PERIODIC]] * 3, max_grid_size=32)
    rlo = np.array([[0, 0, 0], [32, 0, 0]], dtype=np.int32)
    rhi = np.array([[31, 63, 63], [63, 63, 31]], dtype=np.int32)       # second region too small
    with pytest.raises(ValueError):
        PAR.comm_plan(geom, 0, 2, rlo, rhi)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    geom = Geom(3, [32, 32, 64], [[PERIODIC, PERIODIC], [PERIODIC, PERIODIC], [NO_SLIP_WALL, NO_SLIP_WALL]], max_grid_size=32)
    ids, rlo, rhi, pg = PAR.partition(geom, world)
    nbr, g, pc = PAR.comm_plan(geom, rank, world, rlo, rhi)
    plans = [None] * world
    dist.all_gather_object(plans, (nbr.tolist(), pc.tolist(), ids[rank]))
    ok = True
    for d in range(3):
        for s in range(2):
            o = nbr[d][s]
            if o >= 0 and o != rank:
                # my hi neighbour must see me as its lo neighbour (and vice versa)
                ok &= plans[o][0][d][1 - s] == rank
    # the unique id plumbing: rank 0's bytes arrive everywhere
    obj = [os.urandom(128) if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    ids_all = [None] * world
    dist.all_gather_object(ids_all, obj[0])
    ok &= all(x == ids_all[0] for x in ids_all)
    if rank == 0:
        q.put(bool(ok) and sorted(i for p in plans for i in p[2]) == list(range(geom.nboxes)))
    dist.destroy_process_group()


def test_two_rank_plan_is_consistent_over_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
