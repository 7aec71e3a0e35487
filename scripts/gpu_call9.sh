#!/bin/bash
# GPU call 9: k_sweep3 (column pairs per thread, register-pipelined operator data) parity, A/B benches, ncu
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "mg_fused" > gpurun_out/pytest_gpu9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu9.log
tail -5 gpurun_out/pytest_gpu9.log
B="timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e"
for T in 0 1 2; do VDN_MG_FUSE=4 VDN_MG_TILE=$T $B > gpurun_out/b9_f4_tile$T.json 2>&1; done
$B > gpurun_out/b9_f2.json 2>&1
VDN_MG_FUSE=4 VDN_MG_TILE=0 timeout 400 ncu --set full --clock-control none -k regex:'k_sweep3' -s 8 -c 8 -o /tmp/prof_sweep9 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_sweep9.log 2>&1
ncu -i /tmp/prof_sweep9.ncu-rep --page raw --csv > gpurun_out/prof_sweep9_raw.csv 2>/dev/null
ls -la gpurun_out
