#!/bin/bash
# round 2, GPU call 14 (4 GPUs): process grid 2 x 2 x 1 with inflow / outflow in x (the halo-plan fix for a Dirichlet face on a split direction next
# to another split direction), epoch flags pushed to the neighbours, N = 4 strong-scaling point
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
make -C varden_b200/csrc -j16 > gpurun_out/r2c14_build.log 2>&1 || { tail -20 gpurun_out/r2c14_build.log; exit 1; }
timeout 500 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "plain-4-randx3d or fused-4-randx3d or fused_pushk-4-randx3d" > gpurun_out/r2c14_pytest_mgpu4.log 2>&1; tail -4 gpurun_out/r2c14_pytest_mgpu4.log
grep -a "mgpu \|FAILED\|VdnError" gpurun_out/r2c14_pytest_mgpu4.log | head -20
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus 4 --steps 4 --warmup 3 --no-cpu --no-e2e --xchg push > gpurun_out/r2c14_strong_n4_push.json 2> gpurun_out/r2c14_strong_n4_push.err
for f in gpurun_out/r2c14*.err; do echo "== $f"; tail -n 3 "$f"; done
