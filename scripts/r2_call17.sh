#!/bin/bash
# round 2, GPU call 17 (1 GPU): ncu launch list of one 256^3 step (graph nodes are profiled one by one: 515 launches per step), the residual
# history of config 5 at 512^3 (1000:1), GPU tests added since call 16 (estdt golden)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
make -C varden_b200/csrc -j16 > gpurun_out/r2c17_build.log 2>&1 || { tail -20 gpurun_out/r2c17_build.log; exit 1; }
NCU=/usr/local/cuda/bin/ncu
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none --launch-skip 1545 -c 520 --csv --log-file gpurun_out/r2c17_launches.csv \
    python bench.py --config 2 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c17_ncu_launches.log 2>&1
timeout 120 python -m pytest tests/test_golden.py -m gpu -q -k estdt > gpurun_out/r2c17_pytest_estdt.log 2>&1; tail -2 gpurun_out/r2c17_pytest_estdt.log
timeout 400 python scripts/config5_probe.py 512 1000 45 > gpurun_out/r2c17_config5_512.log 2>&1; tail -8 gpurun_out/r2c17_config5_512.log
timeout 200 python scripts/config5_probe.py 256 1000 45 > gpurun_out/r2c17_config5_256.log 2>&1; tail -3 gpurun_out/r2c17_config5_256.log
