#!/bin/bash
# round 2, GPU call 5 (1 GPU): full GPU test tier incl. the new parity tests, single-CTA V-cycle tail, sanitizer pass
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
make -C varden_b200/csrc -j16 > gpurun_out/r2c5_build.log 2>&1 || { tail -20 gpurun_out/r2c5_build.log; exit 1; }
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c5_pytest.log
tail -15 gpurun_out/r2c5_pytest.log
timeout 300 python bench.py --config 2 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c5_256.json 2> gpurun_out/r2c5_256.err
timeout 300 python bench.py --config 2 --ratio 1000 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c5_256_ratio1000.json 2> gpurun_out/r2c5_256_r1000.err
bash scripts/sanitize.sh > gpurun_out/r2c5_sanitize.log 2>&1
tail -12 gpurun_out/r2c5_sanitize.log
