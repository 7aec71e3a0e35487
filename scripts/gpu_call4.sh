#!/bin/bash
# GPU call 4: parity tests, default bench, A/B benches of the fused paths, ncu launch list + full capture (CSV export on the box)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu4.log
tail -5 gpurun_out/pytest_gpu4.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/b4_default.json 2> gpurun_out/b4_default.err
B="timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e"
VDN_MG_FUSE=0 $B > gpurun_out/b4_mgfuse0.json 2>&1
VDN_GODUNOV_FUSE=0 $B > gpurun_out/b4_godfuse0.json 2>&1
VDN_MG_TILE=0 $B > gpurun_out/b4_tile0.json 2>&1
VDN_MG_TILE=1 $B > gpurun_out/b4_tile1.json 2>&1
VDN_MG_FUSE_MIN=64 $B > gpurun_out/b4_min64.json 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r01d.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launches4.log 2>&1
timeout 500 ncu --set full --clock-control none -k regex:'k_wave|k_update|k_mf_|k_vp_|k_mkvelforce|k_wrap|k_absmax_box' -s 60 -c 40 -o /tmp/prof_top4 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_top4.log 2>&1
ncu -i /tmp/prof_top4.ncu-rep --page raw --csv > gpurun_out/prof_top4_raw.csv 2>/dev/null
timeout 300 python bench.py --n 512 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b4_n512.json 2>&1
ls -la gpurun_out
