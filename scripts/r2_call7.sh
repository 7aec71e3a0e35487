#!/bin/bash
# round 2, GPU call 7 (8 GPUs): 8-rank parity (three split directions, corner neighbours) + strong / weak scaling points at N = 8
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
make -C varden_b200/csrc -j16 > gpurun_out/r2c7_build.log 2>&1 || { tail -20 gpurun_out/r2c7_build.log; exit 1; }
nvidia-smi topo -m > gpurun_out/r2c7_topo.txt 2>&1
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "8-rt3d or (fused-8-per3d) or (fused-8-rand3d)" > gpurun_out/r2c7_pytest_mgpu8.log 2>&1; tail -4 gpurun_out/r2c7_pytest_mgpu8.log
T="timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus 8 --config 3 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c7_strong_n8_p2p.json 2> gpurun_out/r2c7_strong_n8_p2p.err
$T bench.py --gpus 8 --config 3 --steps 5 --warmup 3 --no-cpu --no-e2e --force-nccl > gpurun_out/r2c7_strong_n8_nccl.json 2> gpurun_out/r2c7_strong_n8_nccl.err
$T bench.py --gpus 8 --config 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2c7_weak_n8_p2p.json 2> gpurun_out/r2c7_weak_n8_p2p.err
T4="timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512"
$T4 bench.py --gpus 4 --config 3 --steps 4 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c7_strong_n4_p2p.json 2> gpurun_out/r2c7_strong_n4_p2p.err
tail -3 gpurun_out/r2c7*.err
