#!/bin/bash
# GPU call 14 (4 GPUs): single-phase halo exchange (faces + edges in one NCCL group) vs the cascade
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571"
VDN_MG_FUSE_MIN=16 timeout 120 $TR tests/mgpu_worker.py --case rt3d --size 64 > gpurun_out/mgpu14_cascade_rt3d.log 2>&1; echo "rc=$?" >> gpurun_out/mgpu14_cascade_rt3d.log
VDN_HALO_ONEPHASE=1 VDN_MG_FUSE_MIN=16 timeout 120 $TR tests/mgpu_worker.py --case rt3d --size 64 > gpurun_out/mgpu14_onephase_rt3d.log 2>&1; echo "rc=$?" >> gpurun_out/mgpu14_onephase_rt3d.log
VDN_HALO_ONEPHASE=1 VDN_MG_FUSE_MIN=16 timeout 120 $TR tests/mgpu_worker.py --case per3d --size 32 > gpurun_out/mgpu14_onephase_per3d.log 2>&1; echo "rc=$?" >> gpurun_out/mgpu14_onephase_per3d.log
VDN_HALO_ONEPHASE=1 timeout 200 $TR bench.py --gpus 4 --steps 3 --warmup 3 --no-e2e > gpurun_out/b14_n4_onephase.json 2> gpurun_out/b14_n4.err
ls gpurun_out | head -3
