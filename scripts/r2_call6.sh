#!/bin/bash
# round 2, GPU call 6 (2 GPUs): fused smoother reading the neighbour ranks' cells from peer memory (no exchange launch on the fused levels)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
make -C varden_b200/csrc -j16 > gpurun_out/r2c6_build.log 2>&1 || { tail -20 gpurun_out/r2c6_build.log; exit 1; }
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "not (world4 or world8 or 4- or 8-)" > gpurun_out/r2c6_pytest_mgpu.log 2>&1; tail -3 gpurun_out/r2c6_pytest_mgpu.log
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_gpu_parity2.py -m gpu -x -q > gpurun_out/r2c6_pytest_parity2.log 2>&1 &
CUDA_VISIBLE_DEVICES=1 timeout 300 python bench.py --config 2 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c6_256.json 2> gpurun_out/r2c6_256.err
wait
tail -3 gpurun_out/r2c6_pytest_parity2.log
T="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus 2 --config 2 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c6_weak_n2_p2p.json 2> gpurun_out/r2c6_weak_n2_p2p.err
$T bench.py --gpus 2 --config 3 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c6_strong_n2_p2p.json 2> gpurun_out/r2c6_strong_n2_p2p.err
tail -3 gpurun_out/r2c6*.err
