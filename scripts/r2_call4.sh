#!/bin/bash
# round 2, GPU call 4 (2 GPUs): march variants (pair barriers, 2 CTAs/SM, 12-row tile) + fixed pull kernel at N=2
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
make -C varden_b200/csrc -j16 > gpurun_out/r2c4_build.log 2>&1 || { tail -20 gpurun_out/r2c4_build.log; exit 1; }
( cd varden_b200/csrc && for v in "16 1 1" "8 2 0" "8 2 1" "12 1 0" "12 1 1"; do set -- $v
    ( nvcc -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -fmad=false -DMARCH_TYT=$1 -DMARCH_MINB=$2 -DMARCH_PAIRBAR=$3 -c vdn_godunov.cu -o /tmp/god_$1_$2_$3.o &&
      nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libvdn_t$1_m$2_p$3.so vdn_ctx.o /tmp/god_$1_$2_$3.o vdn_stream.o vdn_mg.o vdn_comm.o -lcudart -lnccl ) &
  done; wait )
B="timeout 200 python bench.py --config 2 --steps 5 --warmup 3 --no-cpu --no-e2e"
for v in t16_m1_p1 t8_m2_p0 t8_m2_p1 t12_m1_p0 t12_m1_p1; do
  CUDA_VISIBLE_DEVICES=0 VDN_LIB=$PWD/varden_b200/libvdn_$v.so $B > gpurun_out/r2c4_$v.json 2>&1 &
  sleep 1
  CUDA_VISIBLE_DEVICES=1 VDN_LIB=$PWD/varden_b200/libvdn_$v.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stagewise and 3d" > gpurun_out/r2c4_parity_$v.log 2>&1
  wait
done
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "not (world4 or world8 or 4- or 8-)" > gpurun_out/r2c4_pytest_mgpu.log 2>&1; tail -3 gpurun_out/r2c4_pytest_mgpu.log
T="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus 2 --config 2 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c4_weak_n2_p2p.json 2> gpurun_out/r2c4_weak_n2_p2p.err
$T bench.py --gpus 2 --config 3 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c4_strong_n2_p2p.json 2> gpurun_out/r2c4_strong_n2_p2p.err
tail -3 gpurun_out/*.err
