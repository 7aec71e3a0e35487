#!/bin/bash
# round 2, GPU call 13 (8 GPUs): 8-rank parity (three split directions, corner neighbours, inflow / outflow x) in push / push-kernel modes and
# the north-star point: 512^3 strong scaling at N = 8 (push with e2e, push-kernel A/B)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
make -C varden_b200/csrc -j16 > gpurun_out/r2c13_build.log 2>&1 || { tail -20 gpurun_out/r2c13_build.log; exit 1; }
timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "fused-8-rand3d or fused-8-per3d or fused_pushk-8-rand3d" > gpurun_out/r2c13_pytest_mgpu8.log 2>&1; tail -4 gpurun_out/r2c13_pytest_mgpu8.log
grep -a "mgpu \|FAILED\|Error" gpurun_out/r2c13_pytest_mgpu8.log | head -20
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu --xchg push > gpurun_out/r2c13_strong_n8_push.json 2> gpurun_out/r2c13_strong_n8_push.err
$T bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu --no-e2e --xchg pushk > gpurun_out/r2c13_strong_n8_pushk.json 2> gpurun_out/r2c13_strong_n8_pushk.err
for f in gpurun_out/r2c13*.err; do echo "== $f"; tail -n 3 "$f"; done
