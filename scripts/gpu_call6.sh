#!/bin/bash
# GPU call 6: advance_host parity + e2e bench, ncu full capture of the Godunov stage kernels
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "advance_host or end_to_end or golden" > gpurun_out/pytest_gpu6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu6.log
tail -5 gpurun_out/pytest_gpu6.log
VDN_MG_TILE=0 timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/b6_default.json 2> gpurun_out/b6_default.err
timeout 500 ncu --set full --clock-control none -k regex:'k_mf_|k_vp_|k_update|k_absmax_box|k_mkvelforce' -c 14 -o /tmp/prof_god6 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_god6.log 2>&1
ncu -i /tmp/prof_god6.ncu-rep --page raw --csv > gpurun_out/prof_god6_raw.csv 2>/dev/null
ls -la gpurun_out
