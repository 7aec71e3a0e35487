#!/bin/bash
# round 2, GPU call 8 (2 GPUs): fused smoother in PUSH mode (boundary results stored straight into the neighbours' ghost layers; no exchange
# launches on the fused levels), 2-rank parity, weak / strong N=2 points, and the default single-GPU bench line (512^3, e2e, cpu baseline)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
make -C varden_b200/csrc -j16 > gpurun_out/r2c8_build.log 2>&1 || { tail -20 gpurun_out/r2c8_build.log; exit 1; }
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "not (world4 or world8 or 4- or 8-)" > gpurun_out/r2c8_pytest_mgpu.log 2>&1; tail -3 gpurun_out/r2c8_pytest_mgpu.log
T="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus 2 --config 2 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c8_weak_n2_push.json 2> gpurun_out/r2c8_weak_n2_push.err
$T bench.py --gpus 2 --config 3 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c8_strong_n2_push.json 2> gpurun_out/r2c8_strong_n2_push.err
# single-GPU: the driver's default command on GPU 0, the 256^3 line on GPU 1
( time CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py > gpurun_out/r2c8_default_n1.json 2> gpurun_out/r2c8_default_n1.err ) 2> gpurun_out/r2c8_default_n1.time &
CUDA_VISIBLE_DEVICES=1 timeout 300 python bench.py --config 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2c8_256.json 2> gpurun_out/r2c8_256.err
wait
for f in gpurun_out/r2c8*.err; do echo "== $f"; tail -n 3 "$f"; done
cat gpurun_out/r2c8_default_n1.time
