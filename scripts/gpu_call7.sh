#!/bin/bash
# GPU call 7 (2 GPUs): full parity suite incl. 2-rank cases, fused smoother on rank-split levels, 2-GPU weak-scaling bench
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu7.log
tail -5 gpurun_out/pytest_gpu7.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for c in "rt3d 64" "per3d 32" "rand3d 32"; do set -- $c
  VDN_MG_FUSE_MIN=16 timeout 300 $TR tests/mgpu_worker.py --case $1 --size $2 > gpurun_out/mgpu7_fused_$1.log 2>&1; echo "rc=$?" >> gpurun_out/mgpu7_fused_$1.log; tail -3 gpurun_out/mgpu7_fused_$1.log
done
timeout 400 $TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/b7_n2.json 2> gpurun_out/b7_n2.err; tail -3 gpurun_out/b7_n2.err
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/b7_n1.json 2> gpurun_out/b7_n1.err
ls -la gpurun_out
