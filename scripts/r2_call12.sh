#!/bin/bash
# round 2, GPU call 12 (2 GPUs): a process grid that splits x (2 x 1 x 1) with inflow / outflow boundaries in x -- what the 8-rank rand3d case adds
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
make -C varden_b200/csrc -j16 > gpurun_out/r2c12_build.log 2>&1 || { tail -20 gpurun_out/r2c12_build.log; exit 1; }
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "2-randx3d" > gpurun_out/r2c12_pytest_mgpu.log 2>&1; tail -5 gpurun_out/r2c12_pytest_mgpu.log
grep -a "mgpu " gpurun_out/r2c12_pytest_mgpu.log | head -20
