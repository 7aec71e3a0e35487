#!/bin/bash
# round 2, GPU call 3 (2 GPUs): marching kernels v2 (state in shared-memory keep slots) + pull kernel v2 (flattened grid)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
make -C varden_b200/csrc -j16 > gpurun_out/r2c3_build.log 2>&1 || { tail -20 gpurun_out/r2c3_build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_mock_driver.py -m gpu -x -q > gpurun_out/r2c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c3_pytest.log
tail -6 gpurun_out/r2c3_pytest.log
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --config 2 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c3_256.json 2> gpurun_out/r2c3_256.err
T="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "rt3d and not (world4 or world8 or 4- or 8-)" > gpurun_out/r2c3_pytest_mgpu.log 2>&1; tail -3 gpurun_out/r2c3_pytest_mgpu.log
$T bench.py --gpus 2 --config 2 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c3_weak_n2_p2p.json 2> gpurun_out/r2c3_weak_n2_p2p.err
$T bench.py --gpus 2 --config 2 --steps 5 --warmup 3 --no-cpu --no-e2e --force-nccl > gpurun_out/r2c3_weak_n2_nccl.json 2> gpurun_out/r2c3_weak_n2_nccl.err
CUDA_VISIBLE_DEVICES=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:'march' -s 5 -c 6 -o /tmp/prof_march python bench.py --config 2 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c3_ncu.log 2>&1
cp /tmp/prof_march.ncu-rep gpurun_out/r2c3_march.ncu-rep
ncu -i /tmp/prof_march.ncu-rep --page raw --csv > gpurun_out/r2c3_march_raw.csv 2>/dev/null
tail -3 gpurun_out/*.err
