#!/bin/bash
# round 2, GPU call 19 (1 GPU, the last GPU-minutes of the round): BASELINE config 5 at its stated size (512^3, 1000:1) with the stagnation rule
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
make -C varden_b200/csrc -j16 > gpurun_out/r2c19_build.log 2>&1 || { tail -20 gpurun_out/r2c19_build.log; exit 1; }
timeout 100 python bench.py --config 5 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c19_512_ratio1000.json 2> gpurun_out/r2c19_512_ratio1000.err
tail -n 2 gpurun_out/r2c19_512_ratio1000.err
