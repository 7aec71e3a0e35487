#!/bin/bash
# round 2, GPU call 9 (2 GPUs): A/B of the three peer-memory exchange styles of the fused multigrid levels (push inside the sweep / push kernel /
# pull kernel) with per-rank kernel tables and flag-wait accounting; staged flat host<->device box copies (default 512^3 e2e line)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
make -C varden_b200/csrc -j16 > gpurun_out/r2c9_build.log 2>&1 || { tail -20 gpurun_out/r2c9_build.log; exit 1; }
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "2-rand3d or fused_pushk-2-rt3d or fused-2-per3d" > gpurun_out/r2c9_pytest_mgpu.log 2>&1; tail -3 gpurun_out/r2c9_pytest_mgpu.log
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for x in push pushk pull; do
  $T bench.py --gpus 2 --config 2 --steps 5 --warmup 3 --no-cpu --no-e2e --xchg $x > gpurun_out/r2c9_weak_n2_$x.json 2> gpurun_out/r2c9_weak_n2_$x.err
done
for x in push pushk; do
  $T bench.py --gpus 2 --config 3 --steps 3 --warmup 3 --no-cpu --no-e2e --xchg $x > gpurun_out/r2c9_strong_n2_$x.json 2> gpurun_out/r2c9_strong_n2_$x.err
done
( time CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --no-cpu > gpurun_out/r2c9_default_n1.json 2> gpurun_out/r2c9_default_n1.err ) 2> gpurun_out/r2c9_default_n1.time &
CUDA_VISIBLE_DEVICES=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q -k "box or golden or host" > gpurun_out/r2c9_pytest_boxes.log 2>&1
tail -3 gpurun_out/r2c9_pytest_boxes.log
wait
for f in gpurun_out/r2c9*.err; do echo "== $f"; tail -n 3 "$f"; done
cat gpurun_out/r2c9_default_n1.time
