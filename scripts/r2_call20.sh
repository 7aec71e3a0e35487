#!/bin/bash
# round 2, GPU call 20 (1 GPU, the last minute of the round's GPU budget): the Helmholtz solves (visc_solve / diff_scalar_solve) against the oracle
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
make -C varden_b200/csrc -j16 > gpurun_out/r2c20_build.log 2>&1 || { tail -20 gpurun_out/r2c20_build.log; exit 1; }
timeout 45 python -m pytest tests/test_zz_gpu_helmholtz.py -m gpu -q -x > gpurun_out/r2c20_pytest_helm.log 2>&1; tail -25 gpurun_out/r2c20_pytest_helm.log
