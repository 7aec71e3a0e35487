#!/bin/bash
# first GPU call of this session: parity tests, A/B bench of the fused multigrid, ncu launch list + full captures
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_fuse2.json 2> gpurun_out/bench_fuse2.err
VDN_MG_FUSE=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_fuse0.json 2>&1
VDN_MG_FUSE=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_fuse1.json 2>&1
VDN_MG_TILE=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_fuse2_tile0.json 2>&1
VDN_MG_TILE=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_fuse2_tile1.json 2>&1
VDN_MG_FUSE_MIN=64 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_fuse2_min64.json 2>&1
VDN_MG_FUSE=1 VDN_MG_TILE=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_fuse1_tile0.json 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wave -s 12 -c 4 -o gpurun_out/prof_wave python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_wave.log 2>&1
ls -la gpurun_out
