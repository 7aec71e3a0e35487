#!/bin/bash
# GPU call 15: final validation of the round-1 build: full parity suite, smoke, default bench (e2e + cpu baseline), coarse-level A/B,
# ncu launch list + full capture of the multigrid and Godunov kernels
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu15.log
tail -3 gpurun_out/pytest_gpu15.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke15.log 2>&1; tail -1 gpurun_out/smoke15.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/b15_default.json 2> gpurun_out/b15_default.err
VDN_MG_FUSE_MIN=64 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/b15_min64.json 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r01f.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launches15.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:'k_sweep3|k_mf_normal3|k_vp_normal3' -s 6 -c 24 -o /tmp/prof_top15 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_top15.log 2>&1
ncu -i /tmp/prof_top15.ncu-rep --page raw --csv > gpurun_out/prof_top15_raw.csv 2>/dev/null
ls gpurun_out | wc -l
