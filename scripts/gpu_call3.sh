#!/bin/bash
# GPU call 3: parity tests, A/B benches, ncu launch list + full captures exported to CSV on the box (reports are too large to bring back)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu3.log
tail -5 gpurun_out/pytest_gpu3.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/b3_default.json 2> gpurun_out/b3_default.err
B="timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e"
VDN_MG_FUSE=0 $B > gpurun_out/b3_mgfuse0.json 2>&1
VDN_GODUNOV_FUSE=0 $B > gpurun_out/b3_godfuse0.json 2>&1
VDN_MG_TILE=0 $B > gpurun_out/b3_tile0.json 2>&1
VDN_MG_TILE=1 $B > gpurun_out/b3_tile1.json 2>&1
VDN_MG_TILE=0 VDN_MG_ZCHUNK=64 $B > gpurun_out/b3_tile0_z64.json 2>&1
VDN_MG_TILE=0 VDN_MG_ZCHUNK=128 $B > gpurun_out/b3_tile0_z128.json 2>&1
VDN_MG_TILE=0 VDN_MG_ZCHUNK=256 $B > gpurun_out/b3_tile0_z256.json 2>&1
VDN_MG_FUSE_MIN=64 $B > gpurun_out/b3_min64.json 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launches3.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'k_wave|k_update|k_mf_normal3|k_mf_trans6|k_mf_final3|k_vp_normal3|k_vp_trans6|k_vp_final3|k_mkvelforce|k_wrap|k_absmax_box' -s 60 -c 40 -o /tmp/prof_top3 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_top3.log 2>&1
ncu -i /tmp/prof_top3.ncu-rep --page raw --csv > gpurun_out/prof_top3_raw.csv 2>/dev/null
VDN_MG_FUSE=0 timeout 900 ncu --set full --clock-control none -k regex:'k_gsrb|k_residual|k_restrict|k_prolong' -s 30 -c 12 -o /tmp/prof_mgplain3 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_mgplain3.log 2>&1
ncu -i /tmp/prof_mgplain3.ncu-rep --page raw --csv > gpurun_out/prof_mgplain3_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_wave' -s 8 -c 2 -o gpurun_out/prof_wave3 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_wave3.log 2>&1
timeout 900 python bench.py --n 512 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b3_n512.json 2>&1
python - <<'P' > gpurun_out/pcie.txt 2>&1
import torch, time
a = torch.empty(1 << 28, dtype=torch.float64).pin_memory(); d = torch.empty_like(a, device="cuda")
for name, f in (("h2d", lambda: d.copy_(a, non_blocking=True)), ("d2h", lambda: a.copy_(d, non_blocking=True))):
    f(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(3): f()
    torch.cuda.synchronize(); print(name, 3 * a.numel() * 8 / (time.perf_counter() - t) / 1e9, "GB/s")
P
du -sh gpurun_out; ls -la gpurun_out
