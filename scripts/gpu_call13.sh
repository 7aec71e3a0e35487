#!/bin/bash
# GPU call 13 (2 GPUs): final multi-rank sanity of the default build + the single-phase halo exchange on a one-direction split
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561"
VDN_MG_FUSE_MIN=16 timeout 200 $TR tests/mgpu_worker.py --case rt3d --size 64 > gpurun_out/mgpu13_fused_rt3d.log 2>&1; echo "rc=$?" >> gpurun_out/mgpu13_fused_rt3d.log; tail -2 gpurun_out/mgpu13_fused_rt3d.log
VDN_HALO_ONEPHASE=1 VDN_MG_FUSE_MIN=16 timeout 200 $TR tests/mgpu_worker.py --case rt3d --size 64 > gpurun_out/mgpu13_onephase_rt3d.log 2>&1; echo "rc=$?" >> gpurun_out/mgpu13_onephase_rt3d.log; tail -2 gpurun_out/mgpu13_onephase_rt3d.log
VDN_HALO_ONEPHASE=1 VDN_MG_FUSE_MIN=16 timeout 200 $TR tests/mgpu_worker.py --case per3d --size 32 > gpurun_out/mgpu13_onephase_per3d.log 2>&1; echo "rc=$?" >> gpurun_out/mgpu13_onephase_per3d.log; tail -2 gpurun_out/mgpu13_onephase_per3d.log
timeout 300 $TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/b13_n2.json 2> gpurun_out/b13_n2.err; tail -2 gpurun_out/b13_n2.err
ls -la gpurun_out
