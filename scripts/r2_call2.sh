#!/bin/bash
# round 2, GPU call 2 (2 GPUs): peer-memory transport -- multi-GPU parity, A/B against the NCCL transport, strong-scaling point N=2
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
make -C varden_b200/csrc -j16 > gpurun_out/r2c2_build.log 2>&1 || { tail -20 gpurun_out/r2c2_build.log; exit 1; }
nvidia-smi topo -m > gpurun_out/r2c2_topo.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_gpu_parity.py -m gpu -x -q -k "not (world4 or world8 or 4- or 8-)" > gpurun_out/r2c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c2_pytest.log
tail -15 gpurun_out/r2c2_pytest.log
T="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus 2 --config 2 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c2_weak_n2_p2p.json 2> gpurun_out/r2c2_weak_n2_p2p.err
$T bench.py --gpus 2 --config 2 --steps 5 --warmup 3 --no-cpu --no-e2e --force-nccl > gpurun_out/r2c2_weak_n2_nccl.json 2> gpurun_out/r2c2_weak_n2_nccl.err
$T bench.py --gpus 2 --config 3 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c2_strong_n2_p2p.json 2> gpurun_out/r2c2_strong_n2_p2p.err
tail -3 gpurun_out/*.err
ls -la gpurun_out | tail
