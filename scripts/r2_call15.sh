#!/bin/bash
# round 2, GPU call 15 (4 GPUs): process grid 2 x 2 x 1, inflow / outflow (Dirichlet) in x: the halo-plan fix, fused (push) and plain smoothers
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
make -C varden_b200/csrc -j16 > gpurun_out/r2c15_build.log 2>&1 || { tail -20 gpurun_out/r2c15_build.log; exit 1; }
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "plain-4-randx3d or fused-4-randx3d" > gpurun_out/r2c15_pytest_mgpu4.log 2>&1; tail -4 gpurun_out/r2c15_pytest_mgpu4.log
grep -a "mgpu \|FAILED\|VdnError" gpurun_out/r2c15_pytest_mgpu4.log | head -20
