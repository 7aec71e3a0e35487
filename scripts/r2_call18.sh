#!/bin/bash
# round 2, GPU call 18 (8 GPUs): the 8-rank case that failed before the halo-plan fix (inflow / outflow x, periodic y, walls z over 2 x 2 x 2
# ranks, fused smoother in push mode) and the north-star point with the final build: 512^3 strong scaling at N = 8 incl. e2e
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
make -C varden_b200/csrc -j16 > gpurun_out/r2c18_build.log 2>&1 || { tail -20 gpurun_out/r2c18_build.log; exit 1; }
timeout 200 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "fused-8-rand3d" > gpurun_out/r2c18_pytest_mgpu8.log 2>&1; tail -3 gpurun_out/r2c18_pytest_mgpu8.log
grep -a "mgpu \|FAILED\|VdnError" gpurun_out/r2c18_pytest_mgpu8.log | head -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2c18_strong_n8.json 2> gpurun_out/r2c18_strong_n8.err
tail -n 3 gpurun_out/r2c18_strong_n8.err
