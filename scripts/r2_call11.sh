#!/bin/bash
# round 2, GPU call 11 (2 GPUs): why is the prolongating sweep of level 0 slower in peer-memory push mode?  three experiments (VDN_DBG bits)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
make -C varden_b200/csrc -j16 > gpurun_out/r2c11_build.log 2>&1 || { tail -20 gpurun_out/r2c11_build.log; exit 1; }
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for d in 1 2 4; do
  VDN_DBG=$d $T bench.py --gpus 2 --config 2 --steps 5 --warmup 3 --no-cpu --no-e2e --xchg push > gpurun_out/r2c11_weak_n2_dbg$d.json 2> gpurun_out/r2c11_weak_n2_dbg$d.err
done
for f in gpurun_out/r2c11*.err; do echo "== $f"; tail -n 3 "$f"; done
