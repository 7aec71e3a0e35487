#!/bin/bash
# round 2, GPU call 10 (2 GPUs): fused sweeps pushing AFTER their march (marching loop = the single-GPU one), estdt on the device
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
make -C varden_b200/csrc -j16 > gpurun_out/r2c10_build.log 2>&1 || { tail -20 gpurun_out/r2c10_build.log; exit 1; }
CUDA_VISIBLE_DEVICES=0 timeout 300 python -m pytest tests/test_gpu_parity2.py -m gpu -x -q -k estdt > gpurun_out/r2c10_pytest_estdt.log 2>&1; tail -3 gpurun_out/r2c10_pytest_estdt.log
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "fused-2-" > gpurun_out/r2c10_pytest_mgpu.log 2>&1; tail -3 gpurun_out/r2c10_pytest_mgpu.log
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus 2 --config 2 --steps 5 --warmup 3 --no-cpu --no-e2e --xchg push > gpurun_out/r2c10_weak_n2_push.json 2> gpurun_out/r2c10_weak_n2_push.err
$T bench.py --gpus 2 --config 3 --steps 3 --warmup 3 --no-cpu --no-e2e --xchg push > gpurun_out/r2c10_strong_n2_push.json 2> gpurun_out/r2c10_strong_n2_push.err
for f in gpurun_out/r2c10*.err; do echo "== $f"; tail -n 3 "$f"; done
