"""Multi-GPU parity (-m gpu, needs >= 2 GPUs on the box; skipped otherwise): torchrun + tests/mgpu_worker.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("case,n", [("rt3d", 64), ("rand3d", 32), ("per3d", 32), ("randx3d", 32), ("rt2d", 64)])
@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("smoother", ["plain", "fused", "fused_nccl", "fused_pull", "fused_pushk"])
def test_multi_gpu_parity(case, n, world, smoother):
    """plain: per-colour kernels with a 1-layer exchange per colour; fused: the fused smoother forced onto the rank-split levels (deep
    single-phase ghost exchange: faces, edges and corners in one message set); both through the peer-memory transport; fused_nccl: the same
    exchanges through the NCCL transport"""
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    if case == "rt2d" and (world == 8 or smoother != "plain"):
        pytest.skip("2-D: 4 boxes, plain kernels only")
    if case == "randx3d" and world == 8:
        pytest.skip("x-first split: 2 and 4 ranks (8 ranks split x in the cubic cases)")
    env = dict(os.environ)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29400 + world), os.path.join(ROOT, "tests", "mgpu_worker.py"), "--case", case, "--size", str(n),
           "--fuse-min", "16" if smoother != "plain" else "128",
           "--comm-mode", {"plain": "0", "fused": "0", "fused_nccl": "1", "fused_pull": "2", "fused_pushk": "3"}[smoother]]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    print(r.stdout[:3000], r.stdout[-2000:], r.stderr[-1500:])
    assert r.returncode == 0
