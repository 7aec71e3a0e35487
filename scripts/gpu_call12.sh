#!/bin/bash
# GPU call 12 (4 GPUs): 4-rank parity (plain and fused smoother on the rank-split levels), 4-GPU weak-scaling bench
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541"
for c in "rt3d 64" "per3d 32"; do set -- $c
  timeout 200 $TR tests/mgpu_worker.py --case $1 --size $2 > gpurun_out/mgpu12_plain_$1.log 2>&1; echo "rc=$?" >> gpurun_out/mgpu12_plain_$1.log; tail -2 gpurun_out/mgpu12_plain_$1.log
  VDN_MG_FUSE_MIN=16 timeout 200 $TR tests/mgpu_worker.py --case $1 --size $2 > gpurun_out/mgpu12_fused_$1.log 2>&1; echo "rc=$?" >> gpurun_out/mgpu12_fused_$1.log; tail -2 gpurun_out/mgpu12_fused_$1.log
done
timeout 300 $TR bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/b12_n4.json 2> gpurun_out/b12_n4.err; tail -2 gpurun_out/b12_n4.err
ls -la gpurun_out
