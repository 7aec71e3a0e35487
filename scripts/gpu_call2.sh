#!/bin/bash
# GPU call 2: parity tests, A/B benches (fused MG v2, fused Godunov stages), ncu launch list + full captures of the top kernels
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu2.log
tail -5 gpurun_out/pytest_gpu2.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/b2_default.json 2> gpurun_out/b2_default.err
B="timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e"
VDN_MG_FUSE=0 $B > gpurun_out/b2_mgfuse0.json 2>&1
VDN_GODUNOV_FUSE=0 $B > gpurun_out/b2_godfuse0.json 2>&1
VDN_MG_TILE=0 $B > gpurun_out/b2_tile0.json 2>&1
VDN_MG_TILE=1 $B > gpurun_out/b2_tile1.json 2>&1
VDN_MG_TILE=0 VDN_MG_ZCHUNK=64 $B > gpurun_out/b2_tile0_z64.json 2>&1
VDN_MG_TILE=0 VDN_MG_ZCHUNK=128 $B > gpurun_out/b2_tile0_z128.json 2>&1
VDN_MG_TILE=0 VDN_MG_ZCHUNK=256 $B > gpurun_out/b2_tile0_z256.json 2>&1
VDN_MG_FUSE_MIN=64 $B > gpurun_out/b2_min64.json 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launches2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_wave|k_update|k_mf_normal3|k_mf_trans6|k_mf_final3|k_vp_normal3|k_vp_trans6|k_vp_final3|k_mkvelforce|k_wrap|k_absmax_box' -s 60 -c 40 -o gpurun_out/prof_top2 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_top2.log 2>&1
VDN_MG_FUSE=0 timeout 900 ncu --set full --clock-control none -k regex:'k_gsrb|k_residual|k_restrict|k_prolong' -s 30 -c 12 -o gpurun_out/prof_mgplain2 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_mgplain2.log 2>&1
timeout 900 python bench.py --n 512 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b2_n512.json 2>&1
ls -la gpurun_out
